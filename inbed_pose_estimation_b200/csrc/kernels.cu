// CUDA kernels (sm_100a) of the SMPLify / SMPL hot path and their host launchers.
//
//   smplify_fit_kernel<S>        one CTA = S samples, the whole two-stage fit in one launch
//   pose_forward_kernel<S>       per-sample half of SMPL.forward (joints, skinning transforms, blend coefficients)
//   pose_backward_kernel<S>      its gradient
//   lbs_vertex_forward_kernel    blend shapes + skinning for all 6890 vertices   (smplx lbs, SURVEY §8a a6,a9,a11)
//   lbs_vertex_backward_kernel   dL/dverts -> dL/dA, dL/dx partial sums
//   quat_rodrigues_{fwd,bwd}     utils/geometry.py:9-45
//   projection_{fwd,bwd}         utils/geometry.py:79-107
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "fit_driver.cuh"
#include "launch.h"

namespace smplb200 {

// ------------------------------------------------------------------------------------------------
// tile kernels
// ------------------------------------------------------------------------------------------------
template <int S, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) smplify_fit_kernel(const __grid_constant__ ModelView M,
                                                               const __grid_constant__ FitParams P) {
    extern __shared__ __align__(16) float sm[];
    fit_tile<S>(M, P, blockIdx.x, sm);
}

template <int S>
__global__ void __launch_bounds__(kPoseThreads) pose_forward_kernel(const __grid_constant__ ModelView M,
                                                                    const __grid_constant__ PoseParams P) {
    extern __shared__ __align__(16) float sm[];
    pose_forward_tile<S>(M, P, blockIdx.x, sm);
}

template <int S>
__global__ void __launch_bounds__(kPoseThreads) pose_backward_kernel(const __grid_constant__ ModelView M,
                                                                     const __grid_constant__ PoseParams P) {
    extern __shared__ __align__(16) float sm[];
    pose_backward_tile<S>(M, P, blockIdx.x, sm);
}

template <int S>
static size_t tile_smem_bytes() { return (size_t)TileLayout<S>::SMEM_FLOATS * sizeof(float); }

template <typename K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#if defined(SMPLB200_PHASE_CLOCKS)
cudaError_t debug_phase_clocks(unsigned long long* out32, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out32) e = cudaMemcpyFromSymbol(out32, g_phase_clocks, sizeof(unsigned long long) * 32);
    if (e == cudaSuccess && reset) {
        unsigned long long z[32] = {0};
        e = cudaMemcpyToSymbol(g_phase_clocks, z, sizeof(z));
    }
    return e;
}
#endif

template <int S, int NT, int MINB>
static cudaError_t launch_fit_variant(const ModelView& M, const FitParams& P, cudaStream_t stream) {
    cudaError_t e = opt_in_smem(smplify_fit_kernel<S, NT, MINB>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    smplify_fit_kernel<S, NT, MINB><<<(P.batch + S - 1) / S, NT, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}

cudaError_t launch_fit(const ModelView& M, const FitParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    static const int variant = [] { const char* v = getenv("SMPLB200_FIT_VARIANT"); return v ? atoi(v) : 0; }();
    if (variant == 1) return launch_fit_variant<8, 192, 2>(M, P, stream);
    if (variant == 2) return launch_fit_variant<8, 256, 2>(M, P, stream);
    if (variant == 3) return launch_fit_variant<16, 384, 1>(M, P, stream);
    // Large batches: 16 samples per CTA amortise the streamed folded basis; small ones: spread over more SMs.
    if (P.batch >= 16 * 64) return launch_fit_variant<16, kFitThreads, 1>(M, P, stream);
    if (P.batch >= 8 * 64) return launch_fit_variant<8, kFitThreads, 1>(M, P, stream);
    return launch_fit_variant<4, kFitThreads, 1>(M, P, stream);
}

cudaError_t launch_pose_forward(const ModelView& M, const PoseParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    constexpr int S = 8;
    cudaError_t e = opt_in_smem(pose_forward_kernel<S>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    pose_forward_kernel<S><<<(P.batch + S - 1) / S, kPoseThreads, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}

cudaError_t launch_pose_backward(const ModelView& M, const PoseParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    constexpr int S = 8;
    cudaError_t e = opt_in_smem(pose_backward_kernel<S>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    pose_backward_kernel<S><<<(P.batch + S - 1) / S, kPoseThreads, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// vertex kernels (CUDA-core version): v_posed = x . basis ; verts = (W . A) [v_posed ; 1]
// ------------------------------------------------------------------------------------------------
constexpr int kTV = 64;      // vertices per CTA tile
constexpr int kVTiles = (kVerts + kTV - 1) / kTV;   // 108

template <int TB>
__global__ void __launch_bounds__(256) lbs_vertex_forward_kernel(const __grid_constant__ ModelView M,
                                                                 const float* __restrict__ x, const float* __restrict__ A,
                                                                 float* __restrict__ verts, float* __restrict__ vposed, int batch) {
    constexpr int SPT = TB / 4;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                 // [kX][TB]
    float* As = sm + kX * TB;       // [TB][288]
    const int tid = threadIdx.x, v0 = blockIdx.x * kTV, b0 = blockIdx.y * TB;
    for (int it = tid; it < TB * kX; it += 256) {
        const int s = it / kX, k = it % kX;
        xs[k * TB + s] = (b0 + s < batch) ? x[(size_t)(b0 + s) * kXPad + k] : 0.f;
    }
    for (int it = tid; it < TB * 288; it += 256)
        As[it] = (b0 + it / 288 < batch) ? A[(size_t)b0 * 288 + it] : 0.f;
    __syncthreads();

    const int vl = tid & (kTV - 1), sg = tid / kTV, v = v0 + vl;
    const bool vok = v < kVerts;
    float acc[SPT][3];
#pragma unroll
    for (int i = 0; i < SPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
    const float* bcol = M.basis + 3 * v;         // padded columns are zero, always in bounds
#pragma unroll 2
    for (int k = 0; k < kX; ++k) {
        const float b0v = bcol[(size_t)k * kColsPad + 0], b1v = bcol[(size_t)k * kColsPad + 1], b2v = bcol[(size_t)k * kColsPad + 2];
        const float4* xr = reinterpret_cast<const float4*>(xs + k * TB + sg * SPT);
#pragma unroll
        for (int q = 0; q < SPT / 4; ++q) {
            const float4 xv = xr[q];
            const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[4 * q + u][0] += xx[u] * b0v; acc[4 * q + u][1] += xx[u] * b1v; acc[4 * q + u][2] += xx[u] * b2v;
            }
        }
    }
    float w[kJoints];
    if (vok) {
        const float4* wr = reinterpret_cast<const float4*>(M.weights + (size_t)v * kJoints);
#pragma unroll
        for (int q = 0; q < 6; ++q) { const float4 t = wr[q]; w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w; }
    } else {
#pragma unroll
        for (int j = 0; j < kJoints; ++j) w[j] = 0.f;
    }
#pragma unroll 1
    for (int i = 0; i < SPT; ++i) {
        const int s = sg * SPT + i, b = b0 + s;
        float T[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = 0.f;
        const float4* Ar = reinterpret_cast<const float4*>(As + s * 288);
#pragma unroll
        for (int j = 0; j < kJoints; ++j) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float4 a = Ar[j * 3 + r];
                T[r * 4 + 0] += w[j] * a.x; T[r * 4 + 1] += w[j] * a.y; T[r * 4 + 2] += w[j] * a.z; T[r * 4 + 3] += w[j] * a.w;
            }
        }
        if (vok && b < batch) {
            const float px = acc[i][0], py = acc[i][1], pz = acc[i][2];
            float* o = verts + (size_t)b * kCols + 3 * v;
            o[0] = T[0] * px + T[1] * py + T[2] * pz + T[3];
            o[1] = T[4] * px + T[5] * py + T[6] * pz + T[7];
            o[2] = T[8] * px + T[9] * py + T[10] * pz + T[11];
            if (vposed) {
                float* vp = vposed + (size_t)b * kCols + 3 * v;
                vp[0] = px; vp[1] = py; vp[2] = pz;
            }
        }
    }
}

cudaError_t launch_vertex_forward(const ModelView& M, const float* x, const float* A, float* verts, float* vposed,
                                  int batch, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    constexpr int TB = 32;
    const size_t smem = (size_t)(kX * TB + TB * 288) * sizeof(float);
    cudaError_t e = opt_in_smem(lbs_vertex_forward_kernel<TB>, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(kVTiles, (batch + TB - 1) / TB);
    lbs_vertex_forward_kernel<TB><<<grid, 256, smem, stream>>>(M, x, A, verts, vposed, batch);
    return cudaGetLastError();
}

// Backward: per (sample, vertex)  T = W.A ; dvp = T^R^T dV ; dT = dV (x) [vp;1] ;
//   dA[s][j][e] += W[v][j] dT[s][v][e]      (reduced over the CTA's vertex range)
//   dx[s][m]    += basis[m][3v+c] dvp[s][v][c]
// Each CTA owns TB samples and 1/nsplit of the vertex tiles; partial sums are written once
// (deterministic, no atomics) and added up by pose_backward_kernel.
constexpr int kBwdTB = 16;
__global__ void __launch_bounds__(256) lbs_vertex_backward_kernel(const __grid_constant__ ModelView M,
                                                                  const float* __restrict__ A, const float* __restrict__ vposed,
                                                                  const float* __restrict__ dverts, float* __restrict__ dA_part,
                                                                  float* __restrict__ dx_part, int batch, int nsplit) {
    constexpr int TB = kBwdTB, SPT = TB / 4;
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                          // [TB][288]
    float* Ws = As + TB * 288;               // [kTV][24]
    float* dTs = Ws + kTV * kJoints;         // [kTV][TB][12]
    float* dvps = dTs + kTV * TB * 12;       // [3*kTV][TB]
    const int tid = threadIdx.x, split = blockIdx.x, b0 = blockIdx.y * TB;
    for (int it = tid; it < TB * 288; it += 256)
        As[it] = (b0 + it / 288 < batch) ? A[(size_t)b0 * 288 + it] : 0.f;
    float accA[kJoints], accX[TB];
#pragma unroll
    for (int j = 0; j < kJoints; ++j) accA[j] = 0.f;
#pragma unroll
    for (int s = 0; s < TB; ++s) accX[s] = 0.f;
    const int vl = tid & (kTV - 1), sg = tid / kTV;
    const int t0 = (kVTiles * split) / nsplit, t1 = (kVTiles * (split + 1)) / nsplit;
    __syncthreads();
    for (int vt = t0; vt < t1; ++vt) {
        const int v = vt * kTV + vl;
        const bool vok = v < kVerts;
        float w[kJoints];
        if (vok) {
            const float4* wr = reinterpret_cast<const float4*>(M.weights + (size_t)v * kJoints);
#pragma unroll
            for (int q = 0; q < 6; ++q) { const float4 t = wr[q]; w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w; }
        } else {
#pragma unroll
            for (int j = 0; j < kJoints; ++j) w[j] = 0.f;
        }
        if (sg == 0) {
#pragma unroll
            for (int j = 0; j < kJoints; ++j) Ws[vl * kJoints + j] = w[j];
        }
#pragma unroll 1
        for (int i = 0; i < SPT; ++i) {
            const int s = sg * SPT + i, b = b0 + s;
            float T[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) T[e] = 0.f;
            const float4* Ar = reinterpret_cast<const float4*>(As + s * 288);
#pragma unroll
            for (int j = 0; j < kJoints; ++j) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 a = Ar[j * 3 + r];
                    T[r * 4 + 0] += w[j] * a.x; T[r * 4 + 1] += w[j] * a.y; T[r * 4 + 2] += w[j] * a.z; T[r * 4 + 3] += w[j] * a.w;
                }
            }
            float dV[3] = {0.f, 0.f, 0.f}, vp[3] = {0.f, 0.f, 0.f};
            if (vok && b < batch) {
                const float* g = dverts + (size_t)b * kCols + 3 * v;
                const float* p = vposed + (size_t)b * kCols + 3 * v;
                dV[0] = g[0]; dV[1] = g[1]; dV[2] = g[2];
                vp[0] = p[0]; vp[1] = p[1]; vp[2] = p[2];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                dvps[(3 * vl + c) * TB + s] = T[0 + c] * dV[0] + T[4 + c] * dV[1] + T[8 + c] * dV[2];
            float* dT = dTs + (vl * TB + s) * 12;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                dT[r * 4 + 0] = dV[r] * vp[0]; dT[r * 4 + 1] = dV[r] * vp[1]; dT[r * 4 + 2] = dV[r] * vp[2]; dT[r * 4 + 3] = dV[r];
            }
        }
        __syncthreads();
        if (tid < TB * 12) {                      // thread = (sample, entry of the 3x4), accumulates over joints
            for (int vv = 0; vv < kTV; ++vv) {
                const float d = dTs[vv * TB * 12 + tid];
                const float4* wr = reinterpret_cast<const float4*>(Ws + vv * kJoints);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    const float4 t = wr[q];
                    accA[4 * q] += t.x * d; accA[4 * q + 1] += t.y * d; accA[4 * q + 2] += t.z * d; accA[4 * q + 3] += t.w * d;
                }
            }
        }
        if (tid < kXPad) {                        // thread = blend coefficient m, accumulates over the tile's columns
            const float* bt = M.basisT + (size_t)(3 * vt * kTV) * kXPad + tid;
#pragma unroll 2
            for (int n = 0; n < 3 * kTV; ++n) {
                const float c = bt[(size_t)n * kXPad];
                const float4* dr = reinterpret_cast<const float4*>(dvps + n * TB);
#pragma unroll
                for (int q = 0; q < TB / 4; ++q) {
                    const float4 d = dr[q];
                    accX[4 * q] += c * d.x; accX[4 * q + 1] += c * d.y; accX[4 * q + 2] += c * d.z; accX[4 * q + 3] += c * d.w;
                }
            }
        }
        __syncthreads();
    }
    if (tid < TB * 12) {
        const int s = tid / 12, e = tid % 12, b = b0 + s;
        if (b < batch) {
            float* o = dA_part + ((size_t)split * batch + b) * 288;
#pragma unroll
            for (int j = 0; j < kJoints; ++j) o[j * 12 + e] = accA[j];
        }
    }
    if (tid < kXPad) {
#pragma unroll
        for (int s = 0; s < TB; ++s)
            if (b0 + s < batch) dx_part[((size_t)split * batch + b0 + s) * kXPad + tid] = accX[s];
    }
}

cudaError_t launch_vertex_backward(const ModelView& M, const float* A, const float* vposed, const float* dverts,
                                   float* dA_part, float* dx_part, int batch, int nsplit, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    const size_t smem = (size_t)(kBwdTB * 288 + kTV * kJoints + kTV * kBwdTB * 12 + 3 * kTV * kBwdTB) * sizeof(float);
    cudaError_t e = opt_in_smem(lbs_vertex_backward_kernel, smem);
    if (e != cudaSuccess) return e;
    dim3 grid(nsplit, (batch + kBwdTB - 1) / kBwdTB);
    lbs_vertex_backward_kernel<<<grid, 256, smem, stream>>>(M, A, vposed, dverts, dA_part, dx_part, batch, nsplit);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// utils/geometry.py:9-45   batch_rodrigues (quaternion form) forward / backward
// ------------------------------------------------------------------------------------------------
__global__ void quat_rodrigues_fwd_kernel(const float* __restrict__ theta, float* __restrict__ rot, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rx = theta[3 * i], ry = theta[3 * i + 1], rz = theta[3 * i + 2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float L = sqrtf(ax * ax + ay * ay + az * az);
    const float nx = rx / L, ny = ry / L, nz = rz / L;
    float sh, ch;
    sincosf(L * 0.5f, &sh, &ch);
    float w = ch, x = sh * nx, y = sh * ny, z = sh * nz;
    const float N = sqrtf(w * w + x * x + y * y + z * z);
    w /= N; x /= N; y /= N; z /= N;
    const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    float* R = rot + 9 * (size_t)i;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

__global__ void quat_rodrigues_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ grot,
                                          float* __restrict__ gtheta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rx = theta[3 * i], ry = theta[3 * i + 1], rz = theta[3 * i + 2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float L = sqrtf(ax * ax + ay * ay + az * az);
    const float nx = rx / L, ny = ry / L, nz = rz / L;
    float sh, ch;
    sincosf(L * 0.5f, &sh, &ch);
    const float qw = ch, qx = sh * nx, qy = sh * ny, qz = sh * nz;
    const float N = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    const float w = qw / N, x = qx / N, y = qy / N, z = qz / N;
    const float* g = grot + 9 * (size_t)i;
    const float dw = 2 * w * (g[0] + g[4] + g[8]) + 2 * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
    const float dx = 2 * x * (g[0] - g[4] - g[8]) + 2 * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
    const float dy = 2 * y * (-g[0] + g[4] - g[8]) + 2 * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
    const float dz = 2 * z * (-g[0] - g[4] + g[8]) + 2 * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
    const float dot = w * dw + x * dx + y * dy + z * dz;        // through q / |q|
    const float dqw = (dw - w * dot) / N, dqx = (dx - x * dot) / N, dqy = (dy - y * dot) / N, dqz = (dz - z * dot) / N;
    const float dh = -sh * dqw + ch * (nx * dqx + ny * dqy + nz * dqz);
    const float dnx = sh * dqx, dny = sh * dqy, dnz = sh * dqz;
    float dL = 0.5f * dh;
    dL -= (dnx * rx + dny * ry + dnz * rz) / (L * L);
    gtheta[3 * i + 0] = dnx / L + dL * ax / L;
    gtheta[3 * i + 1] = dny / L + dL * ay / L;
    gtheta[3 * i + 2] = dnz / L + dL * az / L;
}

cudaError_t launch_quat_rodrigues_fwd(const float* theta, float* rot, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    quat_rodrigues_fwd_kernel<<<(n + 255) / 256, 256, 0, st>>>(theta, rot, n);
    return cudaGetLastError();
}
cudaError_t launch_quat_rodrigues_bwd(const float* theta, const float* grot, float* gtheta, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    quat_rodrigues_bwd_kernel<<<(n + 255) / 256, 256, 0, st>>>(theta, grot, gtheta, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// utils/geometry.py:79-107  perspective_projection forward / backward
// ------------------------------------------------------------------------------------------------
__global__ void projection_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ rot, const float* __restrict__ tr,
                                      const float* __restrict__ focal, int focal_per_batch, const float* __restrict__ cen,
                                      float* __restrict__ out, int batch, int npts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)batch * npts) return;
    const int b = (int)(i / npts);
    const float* R = rot + 9 * (size_t)b;
    const float X = pts[3 * i], Y = pts[3 * i + 1], Z = pts[3 * i + 2];
    const float Px = (R[0] * X + R[1] * Y + R[2] * Z) + tr[3 * b + 0];
    const float Py = (R[3] * X + R[4] * Y + R[5] * Z) + tr[3 * b + 1];
    const float Pz = (R[6] * X + R[7] * Y + R[8] * Z) + tr[3 * b + 2];
    const float f = focal[focal_per_batch ? b : 0];
    out[2 * i + 0] = f * (Px / Pz) + cen[2 * b + 0];
    out[2 * i + 1] = f * (Py / Pz) + cen[2 * b + 1];
}

// one CTA per batch element; d_rot / d_tr are reduced over the points in shared memory
__global__ void __launch_bounds__(128) projection_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ rot,
                                                             const float* __restrict__ tr, const float* __restrict__ focal,
                                                             int focal_per_batch, const float* __restrict__ gout,
                                                             float* __restrict__ gpts, float* __restrict__ grot,
                                                             float* __restrict__ gtr, int npts) {
    __shared__ float red[4][12];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* R = rot + 9 * (size_t)b;
    const float f = focal[focal_per_batch ? b : 0];
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) acc[e] = 0.f;
    for (int n = tid; n < npts; n += 128) {
        const size_t i = (size_t)b * npts + n;
        const float X = pts[3 * i], Y = pts[3 * i + 1], Z = pts[3 * i + 2];
        const float Px = (R[0] * X + R[1] * Y + R[2] * Z) + tr[3 * b + 0];
        const float Py = (R[3] * X + R[4] * Y + R[5] * Z) + tr[3 * b + 1];
        const float Pz = (R[6] * X + R[7] * Y + R[8] * Z) + tr[3 * b + 2];
        const float gu = gout[2 * i] * f, gv = gout[2 * i + 1] * f;
        const float dP[3] = {gu / Pz, gv / Pz, -(gu * (Px / Pz) + gv * (Py / Pz)) / Pz};
        gpts[3 * i + 0] = R[0] * dP[0] + R[3] * dP[1] + R[6] * dP[2];
        gpts[3 * i + 1] = R[1] * dP[0] + R[4] * dP[1] + R[7] * dP[2];
        gpts[3 * i + 2] = R[2] * dP[0] + R[5] * dP[1] + R[8] * dP[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            acc[r * 3 + 0] += dP[r] * X; acc[r * 3 + 1] += dP[r] * Y; acc[r * 3 + 2] += dP[r] * Z;
            acc[9 + r] += dP[r];
        }
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        float v = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][e] = v;
    }
    __syncthreads();
    if (tid < 12) {
        const float v = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
        if (tid < 9) grot[9 * (size_t)b + tid] = v;
        else gtr[3 * (size_t)b + tid - 9] = v;
    }
}

cudaError_t launch_projection_fwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* cen, float* out, int batch, int npts, cudaStream_t st) {
    const size_t n = (size_t)batch * npts;
    if (n == 0) return cudaSuccess;
    projection_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, rot, tr, focal, focal_per_batch, cen, out, batch, npts);
    return cudaGetLastError();
}
cudaError_t launch_projection_bwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* gout, float* gpts, float* grot, float* gtr, int batch, int npts, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    projection_bwd_kernel<<<batch, 128, 0, st>>>(pts, rot, tr, focal, focal_per_batch, gout, gpts, grot, gtr, npts);
    return cudaGetLastError();
}

}  // namespace smplb200
