// Experiment: throughput of the fit kernel's folded forward GEMM inner loop (one CTA per SM, 384 threads) in isolation,
// for several register-tile / FFMA forms.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_probe gemm_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int S = 16, kXPad = 224, kQPad = 704, LDQ = S + 4;

struct Tile4x8 {
    float2 a[4][4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < 4; ++p) a[c][p] = make_float2(0.f, 0.f);
    }
    template <int FORM>
    __device__ __forceinline__ void fma(const float4& w, const float4& v0, const float4& v1) {
        const float2 vp[4] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w)};
        const float ws[4] = {w.x, w.y, w.z, w.w};
        if (FORM == 0) {            // compiler-chosen (scalar-broadcast operand)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int p = 0; p < 4; ++p) a[c][p] = __ffma2_rn(make_float2(ws[c], ws[c]), vp[p], a[c][p]);
        } else if (FORM == 1) {     // explicit duplicated pairs
            float2 wp[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                unsigned long long t;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(t) : "f"(ws[c]));
                wp[c] = *reinterpret_cast<float2*>(&t);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int p = 0; p < 4; ++p) a[c][p] = __ffma2_rn(wp[c], vp[p], a[c][p]);
        } else if (FORM == 2) {     // scalar FFMA
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    a[c][p].x = fmaf(ws[c], vp[p].x, a[c][p].x);
                    a[c][p].y = fmaf(ws[c], vp[p].y, a[c][p].y);
                }
        } else {                    // p outer, c inner (sample pair reused across 4 consecutive FFMA2)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int c = 0; c < 4; ++c) a[c][p] = __ffma2_rn(make_float2(ws[c], ws[c]), vp[p], a[c][p]);
        }
    }
};

template <int FORM>
__global__ void __launch_bounds__(384, 1) probe(const float* __restrict__ Cf, float* out, long long* clk, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* XT = sm;                     // [224][16]
    float* QT = sm + kXPad * S;         // [704][LDQ]
    for (int i = threadIdx.x; i < kXPad * S; i += blockDim.x) XT[i] = 0.001f * (i % 97);
    __syncthreads();
    constexpr int U = 4, NQ4 = kQPad / 4, H = S / 8;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int t = threadIdx.x; t < NQ4 * H; t += blockDim.x) {
            const int cq = t % NQ4, h = t / NQ4;
            Tile4x8 acc;
            acc.clear();
            const float4* cf = reinterpret_cast<const float4*>(Cf) + cq;
            const float* xb = XT + 8 * h;
            float4 c0[U], c1[U], c2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { c0[u] = __ldg(cf + u * NQ4); c1[u] = __ldg(cf + (U + u) * NQ4); c2[u] = c1[u]; }
#pragma unroll 1
            for (int m0 = 0; m0 < kXPad; m0 += U) {
                if (m0 + 2 * U < kXPad) {
#pragma unroll
                    for (int u = 0; u < U; ++u) c2[u] = __ldg(cf + (m0 + 2 * U + u) * NQ4);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4* xr = reinterpret_cast<const float4*>(xb + (m0 + u) * S);
                    const float4 v0 = xr[0], v1 = xr[1];
                    acc.fma<FORM>(c0[u], v0, v1);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { c0[u] = c1[u]; c1[u] = c2[u]; }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4* qo = reinterpret_cast<float4*>(QT + (4 * cq + c) * LDQ + 8 * h);
                qo[0] = make_float4(acc.a[c][0].x, acc.a[c][0].y, acc.a[c][1].x, acc.a[c][1].y);
                qo[1] = make_float4(acc.a[c][2].x, acc.a[c][2].y, acc.a[c][3].x, acc.a[c][3].y);
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = QT[5];
}

// 4x8 tile, x rows prefetched one k-step ahead (register double buffer), Cf groups rotated without register moves
__global__ void __launch_bounds__(384, 1) probe_pipe(const float* __restrict__ Cf, float* out, long long* clk, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* XT = sm;
    float* QT = sm + kXPad * S;
    for (int i = threadIdx.x; i < kXPad * S; i += blockDim.x) XT[i] = 0.001f * (i % 97);
    __syncthreads();
    constexpr int U = 4, NQ4 = kQPad / 4, H = S / 8;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int t = threadIdx.x; t < NQ4 * H; t += blockDim.x) {
            const int cq = t % NQ4, h = t / NQ4;
            Tile4x8 acc;
            acc.clear();
            const float4* cf = reinterpret_cast<const float4*>(Cf) + cq;
            const float* xb = XT + 8 * h;
            float4 ca[U], cb[U], cc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { ca[u] = __ldg(cf + u * NQ4); cb[u] = __ldg(cf + (U + u) * NQ4); }
            float4 v0 = reinterpret_cast<const float4*>(xb)[0], v1 = reinterpret_cast<const float4*>(xb)[1];
            auto group = [&](const float4 (&cur)[U], float4 (&nxt)[U], int m0) {
                if (m0 + 2 * U < kXPad) {
#pragma unroll
                    for (int u = 0; u < U; ++u) nxt[u] = __ldg(cf + (m0 + 2 * U + u) * NQ4);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int mn = (m0 + u + 1 < kXPad) ? m0 + u + 1 : m0 + u;
                    const float4* xr = reinterpret_cast<const float4*>(xb + mn * S);
                    const float4 n0 = xr[0], n1 = xr[1];
                    acc.fma<0>(cur[u], v0, v1);
                    v0 = n0; v1 = n1;
                }
            };
            // 224 = 18 * 12 + 8 rows: three groups per round, each reusing the registers of the group two steps back
#pragma unroll 1
            for (int m0 = 0; m0 + 3 * U <= kXPad - 2 * U; m0 += 3 * U) {
                group(ca, cc, m0);
                group(cb, ca, m0 + U);
                group(cc, cb, m0 + 2 * U);
            }
            group(ca, cc, kXPad - 2 * U);
            group(cb, ca, kXPad - U);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4* qo = reinterpret_cast<float4*>(QT + (4 * cq + c) * LDQ + 8 * h);
                qo[0] = make_float4(acc.a[c][0].x, acc.a[c][0].y, acc.a[c][1].x, acc.a[c][1].y);
                qo[1] = make_float4(acc.a[c][2].x, acc.a[c][2].y, acc.a[c][3].x, acc.a[c][3].y);
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = QT[5];
}

// 2 columns x 16 samples per thread: every Cf element is fetched by exactly one thread (no duplicate L2 traffic)
__global__ void __launch_bounds__(384, 1) probe2x16(const float* __restrict__ Cf, float* out, long long* clk, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* XT = sm;
    float* QT = sm + kXPad * S;
    for (int i = threadIdx.x; i < kXPad * S; i += blockDim.x) XT[i] = 0.001f * (i % 97);
    __syncthreads();
    constexpr int U = 4, NQ2 = kQPad / 2;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        for (int t = threadIdx.x; t < NQ2; t += blockDim.x) {
            float2 acc[2][8];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int p = 0; p < 8; ++p) acc[c][p] = make_float2(0.f, 0.f);
            const float2* cf = reinterpret_cast<const float2*>(Cf) + t;
            float2 c0[U], c1[U], c2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { c0[u] = __ldg(cf + u * NQ2); c1[u] = __ldg(cf + (U + u) * NQ2); c2[u] = c1[u]; }
#pragma unroll 1
            for (int m0 = 0; m0 < kXPad; m0 += U) {
                if (m0 + 2 * U < kXPad) {
#pragma unroll
                    for (int u = 0; u < U; ++u) c2[u] = __ldg(cf + (m0 + 2 * U + u) * NQ2);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4* xr = reinterpret_cast<const float4*>(XT + (m0 + u) * S);
                    const float4 v0 = xr[0], v1 = xr[1], v2 = xr[2], v3 = xr[3];
                    const float2 vp[8] = {make_float2(v0.x, v0.y), make_float2(v0.z, v0.w), make_float2(v1.x, v1.y), make_float2(v1.z, v1.w),
                                          make_float2(v2.x, v2.y), make_float2(v2.z, v2.w), make_float2(v3.x, v3.y), make_float2(v3.z, v3.w)};
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        acc[0][p] = __ffma2_rn(make_float2(c0[u].x, c0[u].x), vp[p], acc[0][p]);
                        acc[1][p] = __ffma2_rn(make_float2(c0[u].y, c0[u].y), vp[p], acc[1][p]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { c0[u] = c1[u]; c1[u] = c2[u]; }
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float4* qo = reinterpret_cast<float4*>(QT + (2 * t + c) * LDQ);
#pragma unroll
                for (int q = 0; q < 4; ++q) qo[q] = make_float4(acc[c][2 * q].x, acc[c][2 * q].y, acc[c][2 * q + 1].x, acc[c][2 * q + 1].y);
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = QT[5];
}

template <int FORM>
void run(const char* name, const float* Cf, float* out, long long* clk) {
    const size_t smem = (size_t)(kXPad * S + kQPad * LDQ) * 4;
    cudaFuncSetAttribute(probe<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 50;
    probe<FORM><<<148, 384, smem>>>(Cf, out, clk, reps);
    probe<FORM><<<148, 384, smem>>>(Cf, out, clk, reps);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i] / reps / 148;
    printf("%-28s %8.0f clk per GEMM (ideal 19712 at 128 FMA/clk/SM) -> %.1f%% of peak   err=%s\n", name, mean, 100.0 * 19712 / mean,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *Cf, *out;
    long long* clk;
    cudaMalloc(&Cf, (size_t)kXPad * kQPad * 4);
    cudaMemset(Cf, 0, (size_t)kXPad * kQPad * 4);
    cudaMalloc(&out, 64);
    cudaMalloc(&clk, 148 * 8);
    run<0>("ffma2 scalar-broadcast", Cf, out, clk);
    run<1>("ffma2 explicit pairs", Cf, out, clk);
    run<2>("scalar ffma", Cf, out, clk);
    run<3>("ffma2 p-outer", Cf, out, clk);
    {
        const size_t smem = (size_t)(kXPad * S + kQPad * LDQ) * 4;
        cudaFuncSetAttribute(probe2x16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int nsm : {148, 16}) {
            probe2x16<<<nsm, 384, smem>>>(Cf, out, clk, 50);
            probe2x16<<<nsm, 384, smem>>>(Cf, out, clk, 50);
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, clk, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < nsm; ++i) mean += (double)h[i] / 50 / nsm;
            printf("2 cols x 16 samples, %3d CTAs  %8.0f clk per GEMM -> %.1f%% of peak  err=%s\n", nsm, mean, 100.0 * 19712 / mean,
                   cudaGetErrorString(cudaGetLastError()));
        }
        cudaFuncSetAttribute(probe_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_pipe<<<148, 384, smem>>>(Cf, out, clk, 50);
        probe_pipe<<<148, 384, smem>>>(Cf, out, clk, 50);
        cudaDeviceSynchronize();
        {
            long long h[148];
            cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
            double mean = 0;
            for (int i = 0; i < 148; ++i) mean += (double)h[i] / 50 / 148;
            printf("4x8 tile, pipelined x, no MOVs  %8.0f clk per GEMM -> %.1f%% of peak  err=%s\n", mean, 100.0 * 19712 / mean, cudaGetErrorString(cudaGetLastError()));
        }
        const size_t smem2 = smem;
        cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        probe<0><<<16, 384, smem2>>>(Cf, out, clk, 50);
        probe<0><<<16, 384, smem2>>>(Cf, out, clk, 50);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < 16; ++i) mean += (double)h[i] / 50 / 16;
        printf("4x8 tile on only 16 CTAs        %8.0f clk per GEMM -> %.1f%% of peak (L2 uncontended)\n", mean, 100.0 * 19712 / mean);
    }
    return 0;
}
