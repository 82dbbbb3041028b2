// Experiment (round-2 sizing): cycles per tcgen05.mma (cta_group::1, M = 128, one K step of 32 bytes) as a function of N, for the A
// operand in TMEM (TS) or in shared memory (SS) and kind::tf32 / kind::f16 (bf16 inputs).  One warp per CTA issues R MMAs back to
// back into one accumulator and waits for the commit; operand contents are irrelevant (zeroed shared memory, whatever TMEM holds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../inbed_pose_estimation_b200/csrc -o mma_rate_probe mma_rate_probe.cu
#include "tc_common.cuh"

#include <stdlib.h>

using namespace smplb200::tc;

constexpr int R = 4096;

template <int KIND>   // 0: tf32, 1: bf16
__host__ __device__ constexpr uint32_t idesc_kind(int m, int n) {
    return (1u << 4) | ((KIND == 0 ? 2u : 1u) << 7) | ((KIND == 0 ? 2u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int KIND, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (TS) {
        if (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    } else {
        if (KIND == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
}

template <int KIND, bool TS, int N>
__global__ void __launch_bounds__(64, 1) rate_kernel(long long* cycles) {
    extern __shared__ float smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    float* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) / 4;
    for (int i = threadIdx.x; i < (16 + 32) * 256; i += 64) smem[i] = 0.f;          // A tile 16 KB, B tile up to 32 KB
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 16384);
        constexpr uint32_t idesc = idesc_kind<KIND>(128, N);
        const long long t0 = clock64();
        if (elect_one()) {
#pragma unroll 8
            for (int r = 0; r < R; ++r)
                mma<KIND, TS>(tm, tm + 256 + 8 * (r & 3), da + (uint64_t)(2 * (r & 3)), db + (uint64_t)(2 * (r & 3)), idesc, r ? 1u : 0u);
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        tc_fence_after();
        if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int KIND, bool TS, int N>
void run(long long* d_cyc) {
    const int smem_bytes = 49 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel<KIND, TS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    long long h[148], mx = 0;
    for (int rep = 0; rep < 2; ++rep) {
        rate_kernel<KIND, TS, N><<<148, 64, smem_bytes>>>(d_cyc);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s %s M=128 N=%3d: %.1f cycles per MMA\n", KIND ? "bf16 (K=16)" : "tf32 (K=8) ", TS ? "A in TMEM" : "A in smem", N, (double)mx / R);
}

template <int KIND, bool TS>
void sweep(long long* d) {
    run<KIND, TS, 16>(d); run<KIND, TS, 32>(d); run<KIND, TS, 64>(d); run<KIND, TS, 128>(d); run<KIND, TS, 256>(d);
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    sweep<0, true>(d); sweep<0, false>(d); sweep<1, true>(d); sweep<1, false>(d);
    return 0;
}
