// Tensor-core LBS vertex forward for sm_100a: tcgen05.mma (kind::tf32, accumulators in TMEM) fed by TMA,
// 3xTF32 operand splitting to keep fp32 accuracy.
//
//   GEMM 1 (blend shapes, smplx lbs a6+a9):  v_posed[b, 3v+c] = sum_m x[b,m] * basis[m, 3v+c]     M=128 samples, N=96, K=224
//   GEMM 2 (skinning, a11):                  T_e[b, v]        = sum_j A[b,j,e] * W[v,j], e<12      M=128 samples, N=32, K=24
//   epilogue (CUDA cores, from TMEM):        verts[b,v,r]     = T[4r..4r+2][b,v] . v_posed[b,v,:] + T[4r+3][b,v]
//
// One persistent CTA per SM walks (sample tile, vertex tile) pairs.  Warp roles: warp 0 = TMA producer,
// warp 1 = MMA issuer (one elected thread), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM lane quarter
// = warp % 4).  3xTF32: every fp32 operand is pre-split into hi = tf32(x) and lo = x - hi; each k-step issues
// hi*hi + lo*hi + hi*lo into the same fp32 TMEM accumulator.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "launch.h"
#include "lbs_tc.h"

namespace smplb200 {

namespace tc {

constexpr int BM = 128;                 // samples per tile (UMMA M)
constexpr int BV = 32;                  // vertices per tile
constexpr int BN = 3 * BV;              // 96 coordinate columns (UMMA N of GEMM 1)
constexpr int BK = 32;                  // tf32 elements per 128-byte swizzle row
constexpr int KB = kXPad / BK;          // 7 k-blocks of GEMM 1
constexpr int NE = 12;                  // entries of the 3x4 skinning transform
constexpr int KJ = 3;                   // k-steps (of 8) covering the 24 joints
constexpr int NVT = (kVerts + BV - 1) / BV;     // 216 vertex tiles
constexpr int STAGES1 = 2, STAGES2 = 2;
constexpr uint32_t X_BYTES = BM * BK * 4;       // 16384
constexpr uint32_t B_BYTES = BN * BK * 4;       // 12288
constexpr uint32_t W_BYTES = BV * BK * 4;       // 4096
constexpr uint32_t STAGE1_BYTES = 2 * X_BYTES + 2 * B_BYTES;   // 57344
constexpr uint32_t STAGE2_BYTES = 2 * X_BYTES;                 // 32768
constexpr uint32_t OFF_RING1 = 0;
constexpr uint32_t OFF_RING2 = OFF_RING1 + STAGES1 * STAGE1_BYTES;    // 114688
constexpr uint32_t OFF_W = OFF_RING2 + STAGES2 * STAGE2_BYTES;        // 180224
constexpr uint32_t OFF_BAR = OFF_W + 2 * W_BYTES;                     // 188416
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;                 // + barriers + alignment slack
constexpr int TMEM_COLS = 512;
constexpr int COL_T = BN;               // T_e accumulators start at column 96: 12 x 32 columns
constexpr int THREADS = 256;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    printf("lbs_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t instr_desc(int m, int n) {
    return (1u << 4)            // D format F32
           | (2u << 7)          // A format TF32
           | (2u << 10)         // B format TF32
           | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);      // K-major A and B (bits 15,16 = 0)
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Barriers {
    uint64_t full1[STAGES1], empty1[STAGES1], full2[STAGES2], empty2[STAGES2];
    uint64_t wfull, wempty, tmem_full, tmem_empty;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1)
lbs_vertex_forward_tc_kernel(const __grid_constant__ TcMaps maps, float* __restrict__ verts, float* __restrict__ vposed,
                             int batch, int num_sample_tiles) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operand tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Barriers* bars = reinterpret_cast<Barriers*>(smem + OFF_BAR);
    const uint32_t s_base = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = num_sample_tiles * NVT;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES1; ++i) { mbar_init(smem_u32(&bars->full1[i]), 1); mbar_init(smem_u32(&bars->empty1[i]), 1); }
        for (int i = 0; i < STAGES2; ++i) { mbar_init(smem_u32(&bars->full2[i]), 1); mbar_init(smem_u32(&bars->empty2[i]), 1); }
        mbar_init(smem_u32(&bars->wfull), 1);
        mbar_init(smem_u32(&bars->wempty), 1);
        mbar_init(smem_u32(&bars->tmem_full), 1);
        mbar_init(smem_u32(&bars->tmem_empty), 4);      // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t st1 = 0, ph1 = 0, st2 = 0, ph2 = 0, phw = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int stile = tile / NVT, vt = tile % NVT;
                mbar_wait(smem_u32(&bars->wempty), phw ^ 1);
                mbar_expect_tx(smem_u32(&bars->wfull), 2 * W_BYTES);
                tma_load_2d(s_base + OFF_W, &maps.w_hi, smem_u32(&bars->wfull), 0, vt * BV);
                tma_load_2d(s_base + OFF_W + W_BYTES, &maps.w_lo, smem_u32(&bars->wfull), 0, vt * BV);
                phw ^= 1;
                for (int s = 0; s < KB; ++s) {
                    mbar_wait(smem_u32(&bars->empty1[st1]), ph1 ^ 1);
                    const uint32_t full = smem_u32(&bars->full1[st1]);
                    const uint32_t base = s_base + OFF_RING1 + st1 * STAGE1_BYTES;
                    mbar_expect_tx(full, STAGE1_BYTES);
                    tma_load_2d(base, &maps.x_hi, full, s * BK, stile * BM);
                    tma_load_2d(base + X_BYTES, &maps.x_lo, full, s * BK, stile * BM);
                    tma_load_2d(base + 2 * X_BYTES, &maps.b_hi, full, s * BK, vt * BN);
                    tma_load_2d(base + 2 * X_BYTES + B_BYTES, &maps.b_lo, full, s * BK, vt * BN);
                    if (++st1 == STAGES1) { st1 = 0; ph1 ^= 1; }
                }
                for (int e = 0; e < NE; ++e) {
                    mbar_wait(smem_u32(&bars->empty2[st2]), ph2 ^ 1);
                    const uint32_t full = smem_u32(&bars->full2[st2]);
                    const uint32_t base = s_base + OFF_RING2 + st2 * STAGE2_BYTES;
                    mbar_expect_tx(full, STAGE2_BYTES);
                    tma_load_3d(base, &maps.ae_hi, full, 0, stile * BM, e);
                    tma_load_3d(base + X_BYTES, &maps.ae_lo, full, 0, stile * BM, e);
                    if (++st2 == STAGES2) { st2 = 0; ph2 ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc1 = instr_desc(BM, BN), idesc2 = instr_desc(BM, BV);
            uint32_t st1 = 0, ph1 = 0, st2 = 0, ph2 = 0, phw = 0, pht = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(smem_u32(&bars->tmem_empty), pht ^ 1);        // epilogue drained the accumulators
                tc_fence_after();
                for (int s = 0; s < KB; ++s) {
                    mbar_wait(smem_u32(&bars->full1[st1]), ph1);
                    tc_fence_after();
                    const uint32_t base = s_base + OFF_RING1 + st1 * STAGE1_BYTES;
                    const uint64_t d_xh = smem_desc(base), d_xl = smem_desc(base + X_BYTES);
                    const uint64_t d_bh = smem_desc(base + 2 * X_BYTES), d_bl = smem_desc(base + 2 * X_BYTES + B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);          // 32 bytes per k-step, in 16-byte units
                        umma_tf32(tmem_base, d_xh + ko, d_bh + ko, idesc1, (s | k) != 0);
                        umma_tf32(tmem_base, d_xl + ko, d_bh + ko, idesc1, 1);
                        umma_tf32(tmem_base, d_xh + ko, d_bl + ko, idesc1, 1);
                    }
                    umma_commit(smem_u32(&bars->empty1[st1]));
                    if (++st1 == STAGES1) { st1 = 0; ph1 ^= 1; }
                }
                mbar_wait(smem_u32(&bars->wfull), phw);
                tc_fence_after();
                phw ^= 1;
                const uint64_t d_wh = smem_desc(s_base + OFF_W), d_wl = smem_desc(s_base + OFF_W + W_BYTES);
                for (int e = 0; e < NE; ++e) {
                    mbar_wait(smem_u32(&bars->full2[st2]), ph2);
                    tc_fence_after();
                    const uint32_t base = s_base + OFF_RING2 + st2 * STAGE2_BYTES;
                    const uint64_t d_ah = smem_desc(base), d_al = smem_desc(base + X_BYTES);
                    const uint32_t tm = tmem_base + COL_T + e * BV;
#pragma unroll
                    for (int k = 0; k < KJ; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32(tm, d_ah + ko, d_wh + ko, idesc2, k != 0);
                        umma_tf32(tm, d_al + ko, d_wh + ko, idesc2, 1);
                        umma_tf32(tm, d_ah + ko, d_wl + ko, idesc2, 1);
                    }
                    umma_commit(smem_u32(&bars->empty2[st2]));
                    if (++st2 == STAGES2) { st2 = 0; ph2 ^= 1; }
                }
                umma_commit(smem_u32(&bars->wempty));
                umma_commit(smem_u32(&bars->tmem_full));
                pht ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> registers -> skinning -> HBM =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t pht = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int stile = tile / NVT, vt = tile % NVT;
            const int b = stile * BM + row;
            mbar_wait(smem_u32(&bars->tmem_full), pht);
            tc_fence_after();
            pht ^= 1;
#pragma unroll 1
            for (int c = 0; c < BV / 8; ++c) {
                float vp[24], T[NE][8];
                tmem_ld8(t_lane + 24 * c, vp);
                tmem_ld8(t_lane + 24 * c + 8, vp + 8);
                tmem_ld8(t_lane + 24 * c + 16, vp + 16);
#pragma unroll
                for (int e = 0; e < NE; ++e) tmem_ld8(t_lane + COL_T + e * BV + 8 * c, T[e]);
                tmem_ld_wait();
                float out[24];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        out[3 * i + r] = T[4 * r][i] * vp[3 * i] + T[4 * r + 1][i] * vp[3 * i + 1] + T[4 * r + 2][i] * vp[3 * i + 2] + T[4 * r + 3][i];
                const int v0 = vt * BV + 8 * c;
                if (b < batch) {
                    float* o = verts + (size_t)b * kCols + 3 * v0;          // 8-byte aligned (82680 = 8 * 10335)
                    float* p = vposed ? vposed + (size_t)b * kCols + 3 * v0 : nullptr;
                    if (v0 + 8 <= kVerts) {
#pragma unroll
                        for (int i = 0; i < 12; ++i) reinterpret_cast<float2*>(o)[i] = make_float2(out[2 * i], out[2 * i + 1]);
                        if (p) {
#pragma unroll
                            for (int i = 0; i < 12; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(vp[2 * i], vp[2 * i + 1]);
                        }
                    } else {
                        for (int i = 0; i < 24; ++i)
                            if (3 * v0 + i < kCols) {
                                o[i] = out[i];
                                if (p) p[i] = vp[i];
                            }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp32 row-major [rows][cols] (optionally x planes), box = 32 columns (128 bytes) x box_rows, SWIZZLE_128B
static bool make_map(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint64_t planes, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const bool three_d = planes > 0;
    cuuint64_t dims[3] = {cols, rows, planes};
    cuuint64_t strides[2] = {cols * sizeof(float), cols * rows * sizeof(float)};
    cuuint32_t box[3] = {32, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, three_d ? 3 : 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool tc_make_constant_maps(TcMaps* maps, const float* basisT_hi, const float* basisT_lo, const float* w_hi, const float* w_lo) {
    return make_map(&maps->b_hi, basisT_hi, kXPad, kColsPad, 0, tc::BN) && make_map(&maps->b_lo, basisT_lo, kXPad, kColsPad, 0, tc::BN) &&
           make_map(&maps->w_hi, w_hi, 32, kTcVertRowsPad, 0, tc::BV) && make_map(&maps->w_lo, w_lo, 32, kTcVertRowsPad, 0, tc::BV);
}

cudaError_t launch_vertex_forward_tc(const TcMaps& constant_maps, const TcOperands& op, float* verts, float* vposed, int batch,
                                     cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    TcMaps maps = constant_maps;
    if (!make_map(&maps.x_hi, op.x_hi, kXPad, (uint64_t)batch, 0, tc::BM) || !make_map(&maps.x_lo, op.x_lo, kXPad, (uint64_t)batch, 0, tc::BM) ||
        !make_map(&maps.ae_hi, op.ae_hi, 32, (uint64_t)batch, tc::NE, tc::BM) ||
        !make_map(&maps.ae_lo, op.ae_lo, 32, (uint64_t)batch, tc::NE, tc::BM))
        return cudaErrorInvalidValue;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e = cudaFuncSetAttribute(tc::lbs_vertex_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    const int stiles = (batch + tc::BM - 1) / tc::BM;
    const int tiles = stiles * tc::NVT;
    const int grid = tiles < sms ? tiles : sms;
    tc::lbs_vertex_forward_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(maps, verts, vposed, batch, stiles);
    return cudaGetLastError();
}

}  // namespace smplb200
