// fp32 FMA peak probe: the roofline denominator of the CUDA-core phases is MEASURED on the box
// (MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers).  Scalar FFMA and Blackwell packed FFMA2.
#include <cuda_runtime.h>
#include "../../include/smplify_b200.h"
#include "tc_common.cuh"

namespace {

template <int NACC>
__global__ void probe_scalar(float* out, float a, float b, int iters) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    const float bb[4] = {b, b + 1.f, b + 2.f, b + 3.f};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(a, bb[i & 3], acc[i]);
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(bb[(i + 1) & 3], a, acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void probe_packed(float* out, float a, float b, int iters) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, (float)i);
    const float2 bb[4] = {make_float2(b, b + 1.f), make_float2(b + 2.f, b + 3.f), make_float2(b + 4.f, b + 5.f),
                          make_float2(b + 6.f, b + 7.f)};
    const float2 aa = make_float2(a, a);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(aa, bb[i & 3], acc[i]);
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(bb[(i + 1) & 3], aa, acc[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int smplb200_probe_fp32_peak(int packed, double* tflops) {
    if (!tflops) return 1;
    cudaDeviceProp p;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 1;
    constexpr int N = 16;
    const int threads = 256, blocks = p.multiProcessorCount * 8, iters = 4096;
    float* out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * blocks * threads) != cudaSuccess) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto launch = [&]() {
        if (packed) probe_packed<N><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters);
        else probe_scalar<N><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters);
    };
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess || ms <= 0.f) return 1;
    const double flops = 2.0 * 2 * (packed ? 2 : 1) * N * (double)iters * blocks * threads * reps;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return 0;
}

// tcgen05 kind::tf32 peak probe: one warp per SM issues back-to-back M=128 N=256 K=8 MMAs (operands in shared memory, one
// fp32 accumulator tile in TMEM; the operand values do not matter for the rate).  This is the tensor-pipe denominator of the
// 3xTF32 LBS kernels - MEASURED_PEAKS.json only has the bf16 figure.
namespace {
using namespace smplb200::tc;
constexpr int kTf32ProbeMmas = 8192;

__global__ void __launch_bounds__(64, 1) probe_tf32_kernel(int* sink) {
    extern __shared__ float probe_smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    float* smem = probe_smem_raw + ((1024u - (smem_u32(probe_smem_raw) & 1023u)) & 1023u) / 4;
    for (int i = threadIdx.x; i < (16 + 32) * 256; i += 64) smem[i] = 0.f;          // A tile 16 KB, B tile 32 KB
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 16384);
        constexpr uint32_t idesc = instr_desc(128, 256);
        if (elect_one()) {
#pragma unroll 8
            for (int r = 0; r < kTf32ProbeMmas; ++r)
                umma_tf32(tm, da + (uint64_t)(2 * (r & 3)), db + (uint64_t)(2 * (r & 3)), idesc, r ? 1u : 0u);
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        tc_fence_after();
        if (threadIdx.x == 0 && sink) sink[blockIdx.x] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<256>(tm); }
}
}  // namespace

extern "C" int smplb200_probe_tf32_peak(double* tflops) {
    if (!tflops) return 1;
    cudaDeviceProp p;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess || p.major < 10) return 1;
    const int smem_bytes = 49 * 1024 + 1024, blocks = p.multiProcessorCount;
    if (cudaFuncSetAttribute(probe_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return 1;
    int* sink = nullptr;
    if (cudaMalloc(&sink, sizeof(int) * blocks) != cudaSuccess) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) probe_tf32_kernel<<<blocks, 64, smem_bytes>>>(sink);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) probe_tf32_kernel<<<blocks, 64, smem_bytes>>>(sink);
    cudaEventRecord(e1);
    const cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess || ms <= 0.f) return 1;
    const double flops = 2.0 * 128 * 256 * 8 * (double)kTf32ProbeMmas * blocks * reps;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return 0;
}
