// Experiment (sizing of the next fit-kernel step): how fast can the constants of one fit iteration be streamed
// from L2 into the shared memory of every CTA when the CTAs of a cluster share the fetch?  Each CTA of a cluster of CS
// loads 1/CS of every stage with a 1-D bulk copy and multicasts it to all CS CTAs (cp.async.bulk ... .multicast::cluster);
// stages are released cluster-wide through remote mbarrier arrives.  One CTA per SM, 148 CTAs, no MMA: pure data movement.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o multicast_probe multicast_probe.cu && ./multicast_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#ifndef STAGES
#define STAGES 4
#endif
constexpr int kStageBytes = 32 * 1024, kStages = STAGES, kThreads = 160;   // warp 0 producer, warps 1-4 consumers
constexpr long long kTimeoutClk = 4000000000LL;                       // ~2 s: trap instead of hanging

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && clock64() - t0 > kTimeoutClk) { printf("mbarrier timeout\n"); __trap(); }
    }
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS, bool READ>
__global__ void __launch_bounds__(kThreads, 1) stream_kernel(const char* src, int chunks, int iters, long long* cycles, float* sink) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    const uint32_t rank = (CS > 1) ? cluster_rank() : 0;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 4 * CS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CS > 1) cluster_sync();
    const long long t0 = clock64();
    const int total = chunks * iters;
    constexpr int kSlice = kStageBytes / CS;
    float acc = 0.f;
    if (warp == 0) {
        if (lane == 0) {
            for (int k = 0; k < total; ++k) {
                const int st = k % kStages, use = k / kStages;
                if (use > 0) mbar_wait(smem_u32(&empty[st]), (use - 1) & 1);
                mbar_expect_tx(smem_u32(&full[st]), kStageBytes);
                const char* g = src + (size_t)(k % chunks) * kStageBytes + rank * kSlice;
                const uint32_t dst = smem_u32(smem + st * kStageBytes + rank * kSlice);
                if (CS > 1) {
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                        ::"r"(dst), "l"(g), "r"(kSlice), "r"(smem_u32(&full[st])), "h"((uint16_t)((1u << CS) - 1)) : "memory");
                } else {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(g), "r"(kSlice), "r"(smem_u32(&full[st])) : "memory");
                }
            }
        }
    } else {
        const int cw = warp - 1;                       // 4 consumer warps, a quarter of the stage each
        for (int k = 0; k < total; ++k) {
            const int st = k % kStages, use = k / kStages;
            mbar_wait(smem_u32(&full[st]), use & 1);
            if (READ) {
                const float4* p = reinterpret_cast<const float4*>(smem + st * kStageBytes + cw * (kStageBytes / 4));
#pragma unroll 4
                for (int i = lane; i < kStageBytes / 4 / 16; i += 32) { const float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
            }
            __syncwarp();
            if (lane < CS) mbar_arrive_remote(smem_u32(&empty[st]), lane);      // this warp is done with the stage in this CTA
        }
    }
    __syncthreads();
    if (CS > 1) cluster_sync();                        // no CTA leaves while a peer may still write to it
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (READ && acc == 12345.678f) sink[0] = acc;
}

template <int CS, bool READ>
void run(const char* d_src, int chunks, int iters, long long* d_cyc, float* d_sink, int grid = 148) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kStages * kStageBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(stream_kernel<CS, READ>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, stream_kernel<CS, READ>, d_src, chunks, iters, d_cyc, d_sink);
        cudaEventRecord(e1);
        if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            printf("cluster %d: %s\n", CS, cudaGetErrorString(err != cudaSuccess ? err : cudaGetLastError()));
            exit(1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long h[148];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = (double)chunks * iters * kStageBytes;
    printf("grid %d stages %d cluster %d read %d: %.3f ms, %.0f cycles per pass of %.2f MB, %.1f B/clk into each SM, %.2f TB/s out of L2 chip-wide\n",
           grid, kStages, CS, (int)READ, ms, (double)mx / iters, chunks * kStageBytes / 1e6, bytes / mx, bytes * grid / CS / (ms * 1e-3) / 1e12);
}

int main() {
    const int chunks = 88, iters = 200;                // 88 x 32 KB = 2.88 MB: hi/lo forward + backward constants of one iteration
    char* d_src; long long* d_cyc; float* d_sink;
    cudaMalloc(&d_src, (size_t)chunks * kStageBytes);
    cudaMemset(d_src, 0, (size_t)chunks * kStageBytes);
    cudaMalloc(&d_cyc, 148 * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    run<1, false>(d_src, chunks, iters, d_cyc, d_sink);
    run<2, false>(d_src, chunks, iters, d_cyc, d_sink);
    run<4, false>(d_src, chunks, iters, d_cyc, d_sink);
    run<1, true>(d_src, chunks, iters, d_cyc, d_sink);
    for (int g : {8, 32, 64, 104}) run<1, false>(d_src, chunks, iters, d_cyc, d_sink, g);
    for (int g : {8, 64}) run<2, false>(d_src, chunks, iters, d_cyc, d_sink, g);
    run<2, true>(d_src, chunks, iters, d_cyc, d_sink);
    run<4, true>(d_src, chunks, iters, d_cyc, d_sink);
    return 0;
}
