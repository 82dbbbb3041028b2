// fp32 FMA peak probe: the roofline denominator of the CUDA-core phases is MEASURED on the box
// (MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers).  Scalar FFMA and Blackwell packed FFMA2.
#include <cuda_runtime.h>
#include "../../include/smplify_b200.h"

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
