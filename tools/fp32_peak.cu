// Measures the achievable fp32 FMA rate of the SIMT pipes on this GPU (roofline denominator for the
// CUDA-core phases): scalar FFMA vs Blackwell packed FFMA2 (fma.rn.f32x2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak tools/fp32_peak.cu && ./fp32_peak
#include <cuda_runtime.h>
#include <stdio.h>

template <int NACC>
__global__ void k_scalar(float* out, float a, float b, int iters) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    float bb[4] = {b, b + 1.f, b + 2.f, b + 3.f};
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
__global__ void k_packed(float* out, float a, float b, int iters) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 bb[4] = {make_float2(b, b + 1.f), make_float2(b + 2.f, b + 3.f), make_float2(b + 4.f, b + 5.f), make_float2(b + 6.f, b + 7.f)};
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

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 4096;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk);
    {
        constexpr int N = 16;
        double ms = time_ms([&] { k_scalar<N><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
        double flops = 2.0 * 2 * N * (double)iters * blocks * threads;
        printf(", \"ffma_scalar_tflops\": %.2f", flops / ms / 1e9);
    }
    {
        constexpr int N = 16;
        double ms = time_ms([&] { k_packed<N><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); });
        double flops = 2.0 * 2 * 2 * N * (double)iters * blocks * threads;
        printf(", \"ffma2_packed_tflops\": %.2f", flops / ms / 1e9);
    }
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
