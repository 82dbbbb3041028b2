// Tensor-core LBS vertex path for sm_100a: tcgen05.mma (kind::tf32, fp32 accumulators in TMEM) fed by TMA, with
// 3xTF32 operand splitting (every fp32 operand = hi + lo, hi = tf32(x); each k-step issues hi*hi + lo*hi + hi*lo)
// to keep fp32 accuracy.  Restates smplx lbs and its gradient:
//
//   forward   blend GEMM   v_posed[b, 3v+c] = sum_m x[b,m] basis[m, 3v+c]                 tc_gemm_kernel<256, false>
//             skinning     T[b,v,e] = sum_j W[v,j] A[b,j,e] ; verts = T [v_posed; 1]       tc_skin_kernel<0>
//   backward  skinning^T   dvp[b,v,c] = sum_r T[b,v,r,c] dverts[b,v,r]                     tc_skin_kernel<1>
//             dx GEMM      dx[b,m] = sum_col dvp[b,col] basis[m,col]                       tc_gemm_kernel<224, true>
//             dA GEMM      dA[b,j,(r,c)] = sum_v W[v,j] dverts[b,v,r] [v_posed[b,v]; 1][c]  tc_dA_kernel
//
// Orientation is chosen per GEMM to minimise the operand bytes streamed from L2 (the binding resource once every
// fp32 operand is doubled into hi/lo): K = 224 GEMMs put the samples on the 128 TMEM lanes, the K = 24 skinning
// GEMMs put the VERTICES on the lanes so that the per-tile M operand is the small W block (24 values per row)
// rather than the 288 transform entries per sample.  v_posed / dvp travel through HBM between the kernels.
//
// The tensor pipe truncates its fp32 accumulator after every MMA, a bias that grows with the accumulation chain;
// chains are therefore bounded (<= 84 MMAs for the forward, <= 168 for the gradients) and longer reductions are
// finished in fp32 registers (dx GEMM) or by summing per-split partials (dA GEMM) with round-to-nearest adds.
//
// All kernels are warp-specialised: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warps 4.. = CUDA-core consumers / operand generators (TMEM lane quarter = warp % 4).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "launch.h"
#include "lbs_tc.h"
#include "tc_common.cuh"

namespace smplb200 {
namespace tc {

constexpr int BM = 128;                  // UMMA M (TMEM lanes)
constexpr int BK = 32;                   // tf32 elements per 128-byte swizzle row = one k-block
constexpr uint32_t TILE128 = BM * BK * 4;   // 16384: a [128 x 32] fp32 operand tile

struct Barrier { uint64_t v; };

// =====================================================================================================================
// 3xTF32 GEMM  D[M x N] = A[M x K] . B[N x K]^T  (all K-major).  SPLITK = false: persistent over (m tile, n tile) items,
// K short (one accumulation chain per item), epilogue TMEM -> smem -> TMA store, accumulators double buffered.
// SPLITK = true: one (m tile, K range) item per CTA, chains of CHAIN k-blocks alternate between two TMEM buffers and are
// added into fp32 registers by the flush warps; the partial result is written with plain stores.
// =====================================================================================================================
// hi = x with the 13 low mantissa bits cleared (what the tensor core reads of an fp32 container), lo = x - hi (exact)
__device__ __forceinline__ void split_trunc(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

struct alignas(64) GemmMaps { CUtensorMap a_hi, a_lo, b_hi, b_lo, out, bh_half, bl_half; };    // *_half: box of NT / 2 rows (cluster multicast)
struct GemmWork {
    int num_items;      // m tiles * n tiles * nsplit
    int n_tiles;        // N tiles per m tile
    int nsplit;         // K splits per (m tile, n tile)
    int total_kb;       // k-blocks of the whole K extent
    int chain;          // k-blocks per accumulation chain
};

template <int NT>
struct GemmSmem {
    static constexpr int STAGES = 2;
    static constexpr uint32_t B_BYTES = NT * BK * 4;
    static constexpr uint32_t STAGE_BYTES = 2 * TILE128 + 2 * B_BYTES;
    static constexpr uint32_t OFF_OUT = STAGES * STAGE_BYTES;          // two [128 x 32] staging tiles for the TMA store
    static constexpr uint32_t OFF_BAR = OFF_OUT + 2 * TILE128;
    static constexpr uint32_t BYTES = OFF_BAR + 256 + 1024;            // + barriers + alignment slack
    static_assert(B_BYTES % 1024 == 0, "operand tiles must stay 1024-byte aligned");
    static_assert((B_BYTES / 2) % 1024 == 0 || NT % 16 != 0, "half tiles (cluster multicast) must stay 1024-byte aligned");
    static_assert(BYTES <= 232448, "shared memory budget");
};
struct GemmBars {
    uint64_t full[2], empty[2], tmem_full[2], tmem_empty[2];
    uint64_t split[2];          // SPLITK: the consumer warps have produced the lo half of the stage's A operand
    uint32_t tmem_base;
};

// CL = 2 (SPLITK only): the two CTAs of a cluster take neighbouring m tiles of the same K range; each fetches half of every
// B (basis) tile and multicasts it to both, so the hi / lo basis - the bulk of the operand bytes - leaves L2 once per pair.
template <int NT, bool SPLITK, int CL = 1>
__global__ void __launch_bounds__(SPLITK ? 384 : 256, 1)
tc_gemm_kernel(const __grid_constant__ GemmMaps maps, const GemmWork work, float* __restrict__ out_direct, int batch) {
    static_assert(CL == 1 || (CL == 2 && SPLITK), "clusters are used by the split-K kernel only");
    using SM = GemmSmem<NT>;
    constexpr int STAGES = SM::STAGES;
    constexpr int CONSUMER_WARPS = SPLITK ? 8 : 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + SM::OFF_BAR);
    const uint32_t s_base = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tmem_full[i]), 1); mbar_init(smem_u32(&bars->tmem_empty[i]), CONSUMER_WARPS); }
        for (int i = 0; i < STAGES; ++i) mbar_init(smem_u32(&bars->split[i]), CONSUMER_WARPS);
        fence_barrier_init();
        prefetch_tmap(&maps.a_hi); if (!SPLITK) prefetch_tmap(&maps.a_lo); prefetch_tmap(&maps.b_hi); prefetch_tmap(&maps.b_lo);
        if (!SPLITK) prefetch_tmap(&maps.out);
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(&bars->tmem_base));
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                  // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;

    auto decode = [&](int item, int& mt, int& nt, int& sp, int& kb0, int& kb1) {
        if (CL > 1) item /= CL;                      // the CTAs of a cluster share the item and take m tiles CL * t + rank
        sp = item % work.nsplit;
        const int t = item / work.nsplit;
        nt = t % work.n_tiles;
        mt = (t / work.n_tiles) * CL + (int)crank;
        kb0 = (int)(((long long)work.total_kb * sp) / work.nsplit);
        kb1 = (int)(((long long)work.total_kb * (sp + 1)) / work.nsplit);
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            Ring<STAGES> r;
            for (int item = blockIdx.x; item < work.num_items; item += gridDim.x) {
                int mt, nt, sp, kb0, kb1;
                decode(item, mt, nt, sp, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(smem_u32(&bars->empty[r.stage]), r.phase ^ 1);
                    const uint32_t full = smem_u32(&bars->full[r.stage]);
                    const uint32_t base = s_base + r.stage * SM::STAGE_BYTES;
                    // SPLITK: the A operand arrives as plain fp32 (maps.a_hi) and is split into hi / lo on chip
                    mbar_expect_tx(full, SPLITK ? SM::STAGE_BYTES - TILE128 : SM::STAGE_BYTES);
                    tma_load_2d(base, &maps.a_hi, full, kb * BK, mt * BM);
                    if (!SPLITK) tma_load_2d(base + TILE128, &maps.a_lo, full, kb * BK, mt * BM);
                    if (CL == 1) {
                        tma_load_2d(base + 2 * TILE128, &maps.b_hi, full, kb * BK, nt * NT);
                        tma_load_2d(base + 2 * TILE128 + SM::B_BYTES, &maps.b_lo, full, kb * BK, nt * NT);
                    } else {
                        // this CTA's half of the B rows, delivered to both CTAs (each full barrier still sees the whole tile)
                        const uint32_t hoff = crank * (SM::B_BYTES / 2);
                        tma_load_2d_mc(base + 2 * TILE128 + hoff, &maps.bh_half, full, kb * BK, nt * NT + (int)crank * (NT / 2), 0x3);
                        tma_load_2d_mc(base + 2 * TILE128 + SM::B_BYTES + hoff, &maps.bl_half, full, kb * BK, nt * NT + (int)crank * (NT / 2), 0x3);
                    }
                    r.advance();
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
        constexpr uint32_t idesc = instr_desc(BM, NT);
        Ring<STAGES> r;
        Ring<2> t;
        for (int item = blockIdx.x; item < work.num_items; item += gridDim.x) {
            int mt, nt, sp, kb0, kb1;
            decode(item, mt, nt, sp, kb0, kb1);
            for (int g0 = kb0; g0 < kb1; g0 += work.chain) {
                const int g1 = (g0 + work.chain < kb1) ? g0 + work.chain : kb1;
                mbar_wait(smem_u32(&bars->tmem_empty[t.stage]), t.phase ^ 1);       // consumers drained this accumulator
                tc_fence_after();
                const uint32_t tm = tmem_base + t.stage * NT;
                for (int kb = g0; kb < g1; ++kb) {
                    mbar_wait(smem_u32(SPLITK ? &bars->split[r.stage] : &bars->full[r.stage]), r.phase);
                    tc_fence_after();
                    const uint32_t base = s_base + r.stage * SM::STAGE_BYTES;
                    const uint64_t d_ah = smem_desc(base), d_al = smem_desc(base + TILE128);
                    const uint64_t d_bh = smem_desc(base + 2 * TILE128), d_bl = smem_desc(base + 2 * TILE128 + SM::B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t ko = (uint64_t)(k * 2);          // 32 bytes per k-step, in 16-byte units
                            umma_tf32(tm, d_ah + ko, d_bh + ko, idesc, (kb != g0 || k != 0) ? 1u : 0u);
                            umma_tf32(tm, d_al + ko, d_bh + ko, idesc, 1);
                            umma_tf32(tm, d_ah + ko, d_bl + ko, idesc, 1);
                        }
                        if (CL == 1) umma_commit(smem_u32(&bars->empty[r.stage]));
                        else umma_commit_mc(smem_u32(&bars->empty[r.stage]), 0x3);     // the stage is free when BOTH CTAs have read it
                    }
                    __syncwarp();
                    r.advance();
                }
                if (elect_one()) umma_commit(smem_u32(&bars->tmem_full[t.stage]));
                __syncwarp();
                t.advance();
            }
        }
    } else if (warp >= 4) {
        // ================= consumers: TMEM -> registers -> HBM =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        Ring<2> t;
        if constexpr (!SPLITK) {
            const int epi_tid = threadIdx.x - 128;
            float* staging = reinterpret_cast<float*>(smem + SM::OFF_OUT);
            uint32_t cc = 0;
            for (int item = blockIdx.x; item < work.num_items; item += gridDim.x) {
                int mt, nt, sp, kb0, kb1;
                decode(item, mt, nt, sp, kb0, kb1);
                mbar_wait(smem_u32(&bars->tmem_full[t.stage]), t.phase);
                tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < NT / 32; ++ch, ++cc) {
                    float v[32];
                    tmem_ld32(tmem_base + lane_bits + t.stage * NT + ch * 32, v);
                    tmem_ld_wait();
                    if (epi_tid == 0) tma_store_wait_read<1>();          // the store that last read this staging tile is done
                    named_bar_sync(1, 128);
                    float* sbuf = staging + (cc & 1) * (TILE128 / 4);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(sbuf + swz128(row, c)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    fence_proxy_async();
                    named_bar_sync(1, 128);
                    if (epi_tid == 0) {
                        tma_store_3d(&maps.out, smem_u32(sbuf), nt * NT + ch * 32, mt * BM, sp);
                        tma_store_commit();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[t.stage]));
                t.advance();
            }
            if (epi_tid == 0) tma_store_wait_all<0>();
        } else {
            constexpr int HALF = NT / 2;                      // columns per flush warp (two warps share a lane quarter)
            static_assert(HALF % 16 == 0, "flush width");
            const int half = (warp - 4) >> 2;
            const int ctid = threadIdx.x - 128;               // 0..255 over the 8 consumer warps
            float acc[HALF];
#pragma unroll
            for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
            Ring<STAGES> r;
            // add one finished accumulation chain (TMEM buffer tq) into the fp32 registers and hand the buffer back
            auto flush = [&](const Ring<2>& tq) {
                mbar_wait(smem_u32(&bars->tmem_full[tq.stage]), tq.phase);
                tc_fence_after();
                const uint32_t ta = tmem_base + lane_bits + tq.stage * NT + half * HALF;
#pragma unroll
                for (int c = 0; c < HALF / 8; ++c) {
                    float v[8];
                    tmem_ld8(ta + 8 * c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[8 * c + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[tq.stage]));
            };
            for (int item = blockIdx.x; item < work.num_items; item += gridDim.x) {       // one item per CTA in practice
                int mt, nt, sp, kb0, kb1;
                decode(item, mt, nt, sp, kb0, kb1);
                bool pending = false;
                Ring<2> prev;
                for (int g0 = kb0; g0 < kb1; g0 += work.chain) {
                    const int g1 = (g0 + work.chain < kb1) ? g0 + work.chain : kb1;
                    for (int kb = g0; kb < g1; ++kb) {
                        // split the k-block's fp32 A tile (128 rows x 32 floats, swizzled as TMA wrote it) element-wise, in
                        // place: hi = the 19 bits the tensor core reads, lo = x - hi (exact) into the stage's second tile
                        mbar_wait(smem_u32(&bars->full[r.stage]), r.phase);
                        float4* hi4 = reinterpret_cast<float4*>(smem + r.stage * SM::STAGE_BYTES);
                        float4* lo4 = reinterpret_cast<float4*>(smem + r.stage * SM::STAGE_BYTES + TILE128);
#pragma unroll
                        for (int i = 0; i < (int)(TILE128 / 16) / 256; ++i) {
                            const float4 x = hi4[i * 256 + ctid];
                            float4 h, l;
                            split_trunc(x.x, h.x, l.x); split_trunc(x.y, h.y, l.y); split_trunc(x.z, h.z, l.z); split_trunc(x.w, h.w, l.w);
                            hi4[i * 256 + ctid] = h;
                            lo4[i * 256 + ctid] = l;
                        }
                        fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core's async proxy
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->split[r.stage]));
                        r.advance();
                        // the previous chain is flushed one k-block late: its MMAs have long finished, nothing waits
                        if (kb == g0 && pending) { flush(prev); pending = false; }
                    }
                    prev = t;
                    pending = true;
                    t.advance();
                }
                if (pending) flush(prev);
                const int b = mt * BM + row;
                if (b < batch) {
                    float4* o = reinterpret_cast<float4*>(out_direct + ((size_t)sp * batch + b) * NT + half * HALF);
#pragma unroll
                    for (int c = 0; c < HALF / 4; ++c) o[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                }
#pragma unroll
                for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                  // no CTA leaves while its peer may still multicast to it or arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// =====================================================================================================================
// Skinning GEMM with the vertices on the TMEM lanes:  T^T[v, (b, e)] = sum_j W[v, j] A[b, e, j]   (M = 128 vertices,
// N = 16 samples x 12 entries = 192, K = 24 padded to 32) and the per-vertex transform applied by the consumer warps.
// A CTA owns one 128-vertex block (W tile resident) and walks over sample chunks g, g + groups, ...
// MODE 0: verts[b][v][r] = T[r][0..2] . vp[b][v] + T[r][3]        MODE 1: dvp[b][v][c] = sum_r T[r][c] dverts[b][v][r]
// =====================================================================================================================
struct alignas(64) SkinMaps { CUtensorMap w_hi, w_lo, ae_hi, ae_lo; };
constexpr int SK_NB = 16;                          // samples per chunk
constexpr int SK_N = SK_NB * 12;                   // 192 accumulator columns
constexpr int SK_STAGES = 3;
constexpr uint32_t SK_AE_BYTES = SK_N * BK * 4;    // 24576
constexpr uint32_t SK_STAGE_BYTES = 2 * SK_AE_BYTES;
constexpr uint32_t SK_OFF_RING = 2 * TILE128;
constexpr uint32_t SK_OFF_BAR = SK_OFF_RING + SK_STAGES * SK_STAGE_BYTES;
constexpr uint32_t SK_SMEM = SK_OFF_BAR + 256 + 1024;
constexpr int SK_VBLOCKS = kTcVertRowsPad / BM;    // 54
struct SkinBars {
    uint64_t wfull, full[SK_STAGES], empty[SK_STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(384, 1)
tc_skin_kernel(const __grid_constant__ SkinMaps maps, const float* __restrict__ in, float* __restrict__ out0, float* __restrict__ out1,
               float* __restrict__ out2, int batch, int groups) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    SkinBars* bars = reinterpret_cast<SkinBars*>(smem + SK_OFF_BAR);
    const uint32_t s_base = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int vb = blockIdx.x % SK_VBLOCKS, g = blockIdx.x / SK_VBLOCKS;
    const int nchunks = (batch + SK_NB - 1) / SK_NB;

    if (warp == 0 && lane == 0) {
        mbar_init(smem_u32(&bars->wfull), 1);
        for (int i = 0; i < SK_STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tmem_full[i]), 1); mbar_init(smem_u32(&bars->tmem_empty[i]), 8); }
        fence_barrier_init();
        prefetch_tmap(&maps.w_hi); prefetch_tmap(&maps.w_lo); prefetch_tmap(&maps.ae_hi); prefetch_tmap(&maps.ae_lo);
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(&bars->tmem_base));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t wfull = smem_u32(&bars->wfull);
            mbar_expect_tx(wfull, 2 * TILE128);
            tma_load_2d(s_base, &maps.w_hi, wfull, 0, vb * BM);
            tma_load_2d(s_base + TILE128, &maps.w_lo, wfull, 0, vb * BM);
            Ring<SK_STAGES> r;
            for (int c = g; c < nchunks; c += groups) {
                mbar_wait(smem_u32(&bars->empty[r.stage]), r.phase ^ 1);
                const uint32_t full = smem_u32(&bars->full[r.stage]);
                const uint32_t base = s_base + SK_OFF_RING + r.stage * SK_STAGE_BYTES;
                mbar_expect_tx(full, SK_STAGE_BYTES);
                tma_load_2d(base, &maps.ae_hi, full, 0, c * SK_N);
                tma_load_2d(base + SK_AE_BYTES, &maps.ae_lo, full, 0, c * SK_N);
                r.advance();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = instr_desc(BM, SK_N);
        const uint64_t d_wh = smem_desc(s_base), d_wl = smem_desc(s_base + TILE128);
        mbar_wait(smem_u32(&bars->wfull), 0);
        tc_fence_after();
        Ring<SK_STAGES> r;
        Ring<2> t;
        for (int c = g; c < nchunks; c += groups) {
            mbar_wait(smem_u32(&bars->tmem_empty[t.stage]), t.phase ^ 1);
            mbar_wait(smem_u32(&bars->full[r.stage]), r.phase);
            tc_fence_after();
            const uint32_t base = s_base + SK_OFF_RING + r.stage * SK_STAGE_BYTES;
            const uint64_t d_ah = smem_desc(base), d_al = smem_desc(base + SK_AE_BYTES);
            const uint32_t tm = tmem_base + t.stage * SK_N;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {                 // 24 joints = 3 k-steps of 8
                    const uint64_t ko = (uint64_t)(k * 2);
                    umma_tf32(tm, d_wh + ko, d_ah + ko, idesc, k != 0);
                    umma_tf32(tm, d_wl + ko, d_ah + ko, idesc, 1);
                    umma_tf32(tm, d_wh + ko, d_al + ko, idesc, 1);
                }
                umma_commit(smem_u32(&bars->empty[r.stage]));
                umma_commit(smem_u32(&bars->tmem_full[t.stage]));
            }
            __syncwarp();
            r.advance();
            t.advance();
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int v = vb * BM + q * 32 + lane;
        const bool vok = v < kVerts;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        Ring<2> t;
        // this thread's inputs of one chunk: 8 samples x 3 coordinates; the next chunk's are fetched before the current
        // one is transformed (software pipeline: keeps HBM reads in flight while the warp waits on TMEM / stores)
        auto fetch = [&](int c, float (&dst)[8][3]) {
            const int b0 = c * SK_NB + half * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = b0 + i;
                if (MODE == 0) {
                    const bool ok = c < nchunks && b < batch;            // pitch covers the padded vertices
                    const float* p = in + (size_t)b * kVpPitch + 3 * v;
#pragma unroll
                    for (int a = 0; a < 3; ++a) dst[i][a] = ok ? __ldg(p + a) : 0.f;
                } else {
                    const bool ok = c < nchunks && vok && b < batch;
                    const float* p = in + (size_t)b * kCols + 3 * v;
#pragma unroll
                    for (int a = 0; a < 3; ++a) dst[i][a] = ok ? __ldg(p + a) : 0.f;
                }
            }
        };
        // two chunks ahead: the kernel is bound by latency x bytes in flight (8 consumer warps per SM), not by issue slots
        float iv[8][3], nx[8][3], nx2[8][3];
        fetch(g, nx);
        fetch(g + groups, nx2);
        for (int c = g; c < nchunks; c += groups) {
            const int b0 = c * SK_NB + half * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int a = 0; a < 3; ++a) { iv[i][a] = nx[i][a]; nx[i][a] = nx2[i][a]; }
            fetch(c + 2 * groups, nx2);
            mbar_wait(smem_u32(&bars->tmem_full[t.stage]), t.phase);
            tc_fence_after();
            const uint32_t ta = tmem_base + lane_bits + t.stage * SK_N + half * 96;
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {                  // two samples (24 columns) per round
                float T[24];
                tmem_ld8(ta + 24 * pr, T);
                tmem_ld8(ta + 24 * pr + 8, T + 8);
                tmem_ld8(ta + 24 * pr + 16, T + 16);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = 2 * pr + u, b = b0 + i;
                    const float* Tt = T + 12 * u;
                    const float a0 = iv[i][0], a1 = iv[i][1], a2 = iv[i][2];
                    if (MODE == 0) {
                        if (vok && b < batch) {
                            float* o = out0 + (size_t)b * kCols + 3 * v;
                            o[0] = Tt[0] * a0 + Tt[1] * a1 + Tt[2] * a2 + Tt[3];
                            o[1] = Tt[4] * a0 + Tt[5] * a1 + Tt[6] * a2 + Tt[7];
                            o[2] = Tt[8] * a0 + Tt[9] * a1 + Tt[10] * a2 + Tt[11];
                        }
                    } else {
                        if (b < batch) {                                 // padded vertices get zeros (dverts masked above)
                            float* od = out0 + (size_t)b * kVpPitch + 3 * v;     // dvp in fp32: the dx GEMM splits it on chip
                            float* oc = out2 + (size_t)b * kVpPitch + 3 * v;     // 16-byte aligned copy of dverts for the dA kernel's TMA
                            oc[0] = a0; oc[1] = a1; oc[2] = a2;
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) od[cc] = Tt[cc] * a0 + Tt[4 + cc] * a1 + Tt[8 + cc] * a2;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[t.stage]));
            t.advance();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// =====================================================================================================================
// dA GEMM with operands generated on chip:  D_e[b, j] = sum_v P_e[b, v] W[v, j],  P_(r,c)[b, v] = dverts[b,v,r] [vp[b,v]; 1][c]
// (M = 128 samples, N = 24 joints padded to 32, K = the split's vertex range in k-blocks of 32; 12 accumulators of 32
// columns).  The TMA producer streams the k-block's dverts / v_posed tiles ([128 samples][96 coordinates], from the
// padded 16-byte aligned copies) through a two-stage ring; generator warps read one row-half each (thread = TMEM lane
// = sample row, 16 of the 32 vertices), form the 12 product tiles, split them into hi / lo and store them straight into
// TENSOR MEMORY, from where the MMA reads its A operand (tcgen05.mma with A in TMEM): no shared-memory round trip and
// no operand-fetch bottleneck for the narrow N = 32 MMAs.
// TMEM columns: 0..383 accumulators, 384..511 two P stages of (hi 32 | lo 32) columns.
// =====================================================================================================================
struct alignas(64) DaMaps { CUtensorMap wT_hi, wT_lo, dv, vp; };
constexpr int DA_PSTAGES = 2, DA_WSTAGES = 2, DA_ISTAGES = 2;
constexpr int DA_P_COL = 384;                                  // first TMEM column of the P stages
constexpr uint32_t DA_W_TILE = 32 * BK * 4;                    // 4096
constexpr uint32_t DA_IN_STAGE = 6 * TILE128;                  // dverts: 3 column blocks of [128 x 32] | v_posed: 3 blocks
constexpr uint32_t DA_OFF_IN = 0;
constexpr uint32_t DA_OFF_W = DA_ISTAGES * DA_IN_STAGE;        // 196608
constexpr uint32_t DA_OFF_BAR = DA_OFF_W + DA_WSTAGES * 2 * DA_W_TILE;
constexpr uint32_t DA_SMEM = DA_OFF_BAR + 256 + 1024;
static_assert(DA_SMEM <= 232448, "shared memory budget");
constexpr int DA_KBLOCKS = kTcVertRowsPad / BK;                // 216
struct DaBars {
    uint64_t wfull[DA_WSTAGES], wempty[DA_WSTAGES], ifull[DA_ISTAGES], iempty[DA_ISTAGES], pfull[DA_PSTAGES], pempty[DA_PSTAGES], acc_full;
    uint32_t tmem_base;
};



__global__ void __launch_bounds__(384, 1)
tc_dA_kernel(const __grid_constant__ DaMaps maps, float* __restrict__ dA_part, int batch, int nsplit) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    DaBars* bars = reinterpret_cast<DaBars*>(smem + DA_OFF_BAR);
    const uint32_t s_base = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sp = blockIdx.x, st = blockIdx.y;
    const int kb0 = (DA_KBLOCKS * sp) / nsplit, kb1 = (DA_KBLOCKS * (sp + 1)) / nsplit;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < DA_WSTAGES; ++i) { mbar_init(smem_u32(&bars->wfull[i]), 1); mbar_init(smem_u32(&bars->wempty[i]), 1); }
        for (int i = 0; i < DA_ISTAGES; ++i) { mbar_init(smem_u32(&bars->ifull[i]), 1); mbar_init(smem_u32(&bars->iempty[i]), 8); }
        for (int i = 0; i < DA_PSTAGES; ++i) { mbar_init(smem_u32(&bars->pfull[i]), 8); mbar_init(smem_u32(&bars->pempty[i]), 1); }
        mbar_init(smem_u32(&bars->acc_full), 1);
        fence_barrier_init();
        prefetch_tmap(&maps.wT_hi); prefetch_tmap(&maps.wT_lo); prefetch_tmap(&maps.dv); prefetch_tmap(&maps.vp);
    }
    if (warp == 2) tmem_alloc<512>(smem_u32(&bars->tmem_base));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ================= TMA producer: W^T tiles and the dverts / v_posed tiles of every k-block =================
        if (lane == 0) {
            Ring<DA_WSTAGES> wr;
            Ring<DA_ISTAGES> ir;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(smem_u32(&bars->iempty[ir.stage]), ir.phase ^ 1);
                const uint32_t ifull = smem_u32(&bars->ifull[ir.stage]);
                const uint32_t ibase = s_base + DA_OFF_IN + ir.stage * DA_IN_STAGE;
                mbar_expect_tx(ifull, DA_IN_STAGE);
#pragma unroll
                for (int cb = 0; cb < 3; ++cb) {
                    tma_load_2d(ibase + cb * TILE128, &maps.dv, ifull, (3 * kb + cb) * BK, st * BM);
                    tma_load_2d(ibase + (3 + cb) * TILE128, &maps.vp, ifull, (3 * kb + cb) * BK, st * BM);
                }
                ir.advance();
                mbar_wait(smem_u32(&bars->wempty[wr.stage]), wr.phase ^ 1);
                const uint32_t wfull = smem_u32(&bars->wfull[wr.stage]);
                const uint32_t wbase = s_base + DA_OFF_W + wr.stage * 2 * DA_W_TILE;
                mbar_expect_tx(wfull, 2 * DA_W_TILE);
                tma_load_2d(wbase, &maps.wT_hi, wfull, kb * BK, 0);
                tma_load_2d(wbase + DA_W_TILE, &maps.wT_lo, wfull, kb * BK, 0);
                wr.advance();
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
        constexpr uint32_t idesc = instr_desc(BM, 32);
        Ring<DA_WSTAGES> wr;
        Ring<DA_PSTAGES> pr;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&bars->wfull[wr.stage]), wr.phase);
            tc_fence_after();
            const uint32_t wbase = s_base + DA_OFF_W + wr.stage * 2 * DA_W_TILE;
            const uint64_t d_wh = smem_desc(wbase), d_wl = smem_desc(wbase + DA_W_TILE);
            const uint32_t first = (kb != kb0) ? 1u : 0u;
#pragma unroll 1
            for (int e = 0; e < 12; ++e) {
                mbar_wait(smem_u32(&bars->pfull[pr.stage]), pr.phase);
                tc_fence_after();
                const uint32_t t_ph = tmem_base + DA_P_COL + pr.stage * 64, t_pl = t_ph + 32;
                const uint32_t tm = tmem_base + e * 32;
                if (elect_one()) {
                    umma_tf32_ts(tm, t_ph, d_wh, idesc, first);
                    umma_tf32_ts(tm, t_pl, d_wh, idesc, 1);
                    umma_tf32_ts(tm, t_ph, d_wl, idesc, 1);
#pragma unroll
                    for (int k = 1; k < BK / 8; ++k) {
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_tf32_ts(tm, t_ph + 8 * k, d_wh + ko, idesc, 1);
                        umma_tf32_ts(tm, t_pl + 8 * k, d_wh + ko, idesc, 1);
                        umma_tf32_ts(tm, t_ph + 8 * k, d_wl + ko, idesc, 1);
                    }
                    umma_commit(smem_u32(&bars->pempty[pr.stage]));
                }
                __syncwarp();
                pr.advance();
            }
            if (elect_one()) umma_commit(smem_u32(&bars->wempty[wr.stage]));
            __syncwarp();
            wr.advance();
        }
        if (elect_one()) umma_commit(smem_u32(&bars->acc_full));
        __syncwarp();
    } else if (warp >= 4) {
        // ================= operand generators (8 warps): thread = (sample row = TMEM lane, half of the k-block's 32 vertices) ======
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int rowl = q * 32 + lane;
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        Ring<DA_PSTAGES> pr;
        Ring<DA_ISTAGES> ir;
        for (int kb = kb0; kb < kb1; ++kb) {
            float dv[48], vp[48];
            mbar_wait(smem_u32(&bars->ifull[ir.stage]), ir.phase);
            {
                const float* tile = reinterpret_cast<const float*>(smem + DA_OFF_IN + ir.stage * DA_IN_STAGE);
#pragma unroll
                for (int i = 0; i < 12; ++i) {                       // this thread's 48 coordinates = 12 16-byte chunks of the 96-column row
                    const int ch = 12 * h + i, cb = ch >> 3, cc = ch & 7;
                    const float4 a = *reinterpret_cast<const float4*>(tile + cb * (TILE128 / 4) + swz128(rowl, cc));
                    const float4 c = *reinterpret_cast<const float4*>(tile + (3 + cb) * (TILE128 / 4) + swz128(rowl, cc));
                    dv[4 * i] = a.x; dv[4 * i + 1] = a.y; dv[4 * i + 2] = a.z; dv[4 * i + 3] = a.w;
                    vp[4 * i] = c.x; vp[4 * i + 1] = c.y; vp[4 * i + 2] = c.z; vp[4 * i + 3] = c.w;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->iempty[ir.stage]));       // the row-halves are in registers: refill the stage
            ir.advance();
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const int r = e >> 2, c = e & 3;
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p = (c < 3) ? dv[3 * i + r] * vp[3 * i + c] : dv[3 * i + r];
                    split_trunc(p, hi[i], lo[i]);
                }
                mbar_wait(smem_u32(&bars->pempty[pr.stage]), pr.phase ^ 1);
                tc_fence_after();
                const uint32_t ta = tmem_base + lane_bits + DA_P_COL + pr.stage * 64 + 16 * h;
                tmem_st16(ta, hi);
                tmem_st16(ta + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->pfull[pr.stage]));
                pr.advance();
            }
        }
        // ================= epilogue (warps 4..7): D_e[b][0..23] -> dA_part[split][b][e][24] =================
        if (warp < 8) {
            mbar_wait(smem_u32(&bars->acc_full), 0);
            tc_fence_after();
            const int bb = st * BM + rowl;
#pragma unroll 1
            for (int e = 0; e < 12; ++e) {
                float v[32];
                tmem_ld32(tmem_base + lane_bits + e * 32, v);
                tmem_ld_wait();
                if (bb < batch) {
                    float4* o = reinterpret_cast<float4*>(dA_part + (((size_t)sp * batch + bb) * 12 + e) * kJoints);
#pragma unroll
                    for (int c = 0; c < 6; ++c) o[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
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

// fp32 row-major [planes][rows][cols] (planes = 0: two-dimensional), box = 32 columns (128 bytes) x box_rows, SWIZZLE_128B
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

bool tc_make_constant_maps(TcConstMaps* m, const float* basisT_hi, const float* basisT_lo, const float* basis_hi,
                           const float* basis_lo, const float* w_hi, const float* w_lo, const float* wT_hi, const float* wT_lo) {
    return make_map(&m->bT_hi, basisT_hi, kXPad, kColsPad, 0, 256) && make_map(&m->bT_lo, basisT_lo, kXPad, kColsPad, 0, 256) &&
           make_map(&m->bm_hi, basis_hi, kColsPad, kXPad, 0, 224) && make_map(&m->bm_lo, basis_lo, kColsPad, kXPad, 0, 224) &&
           make_map(&m->bm_hi_half, basis_hi, kColsPad, kXPad, 0, 112) && make_map(&m->bm_lo_half, basis_lo, kColsPad, kXPad, 0, 112) &&
           make_map(&m->w_hi, w_hi, 32, kTcVertRowsPad, 0, 128) && make_map(&m->w_lo, w_lo, 32, kTcVertRowsPad, 0, 128) &&
           make_map(&m->wT_hi, wT_hi, kTcVertRowsPad, 32, 0, 32) && make_map(&m->wT_lo, wT_lo, kTcVertRowsPad, 32, 0, 32);
}

static int sm_count() { return device_sm_count(); }      // kernels.cu: cached per device

template <typename K>
static cudaError_t opt_in(K kernel, uint32_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

cudaError_t launch_blend_gemm(const TcConstMaps& cm, const float* x_hi, const float* x_lo, float* vposed, int batch, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    constexpr int NT = 256;
    tc::GemmMaps maps;
    maps.b_hi = cm.bT_hi;
    maps.b_lo = cm.bT_lo;
    if (!make_map(&maps.a_hi, x_hi, kXPad, (uint64_t)batch, 0, tc::BM) || !make_map(&maps.a_lo, x_lo, kXPad, (uint64_t)batch, 0, tc::BM) ||
        !make_map(&maps.out, vposed, kVpPitch, (uint64_t)batch, 1, tc::BM))
        return cudaErrorInvalidValue;
    tc::GemmWork w;
    const int mtiles = (batch + tc::BM - 1) / tc::BM;
    w.n_tiles = kColsPad / NT;           // 81
    w.nsplit = 1;
    w.total_kb = kXPad / tc::BK;         // 7
    w.chain = w.total_kb;
    w.num_items = mtiles * w.n_tiles;
    cudaError_t e = opt_in(tc::tc_gemm_kernel<NT, false>, tc::GemmSmem<NT>::BYTES);
    if (e != cudaSuccess) return e;
    const int grid = w.num_items < sm_count() ? w.num_items : sm_count();
    tc::tc_gemm_kernel<NT, false><<<grid, 256, tc::GemmSmem<NT>::BYTES, stream>>>(maps, w, nullptr, batch);
    return cudaGetLastError();
}

// K splits of the dx GEMM: the smallest count whose CTA grid fills whole waves of SMs to >= 90 % (else the best one)
int tc_dx_splits(int batch) {
    const int mtiles = (batch + tc::BM - 1) / tc::BM, sms = sm_count();
    int best = 1;
    double best_eff = 0.0;
    for (int ns = 1; ns <= kMaxSplitX; ++ns) {
        const int ctas = mtiles * ns, waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / ((double)waves * sms);
        if (eff >= 0.9) return ns;
        if (eff > best_eff) { best_eff = eff; best = ns; }
    }
    return best;
}

cudaError_t launch_dx_gemm(const TcConstMaps& cm, const float* dvp, float* dx_part, int batch, int nsplit, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    constexpr int NT = 224;
    tc::GemmMaps maps;
    maps.b_hi = cm.bm_hi;
    maps.b_lo = cm.bm_lo;
    maps.bh_half = cm.bm_hi_half;
    maps.bl_half = cm.bm_lo_half;
    if (!make_map(&maps.a_hi, dvp, kVpPitch, (uint64_t)batch, 0, tc::BM)) return cudaErrorInvalidValue;
    maps.a_lo = maps.a_hi;               // the split-K kernel derives the lo half on chip
    maps.out = maps.a_hi;                // unused by the split-K epilogue
    tc::GemmWork w;
    const int mtiles = (batch + tc::BM - 1) / tc::BM;
    w.n_tiles = 1;
    w.nsplit = nsplit;
    w.total_kb = kVpPitch / tc::BK;      // 648
    w.chain = 7;                          // 84 MMAs per accumulation chain
    static const bool no_cluster = getenv("SMPLB200_DX_NOCLUSTER") != nullptr;          // experiments only
    auto launch_single = [&]() -> cudaError_t {
        w.num_items = mtiles * nsplit;
        cudaError_t e = opt_in(tc::tc_gemm_kernel<NT, true, 1>, tc::GemmSmem<NT>::BYTES);
        if (e != cudaSuccess) return e;
        tc::tc_gemm_kernel<NT, true, 1><<<w.num_items, 384, tc::GemmSmem<NT>::BYTES, stream>>>(maps, w, dx_part, batch);
        return cudaGetLastError();
    };
    if (mtiles < 2 || no_cluster) return launch_single();
    // pairs of m tiles share the basis tiles through a 2-CTA cluster (an odd last tile is paired with an all-padding one)
    const int mpairs = (mtiles + 1) / 2;
    w.num_items = mpairs * nsplit * 2;
    cudaError_t e = opt_in(tc::tc_gemm_kernel<NT, true, 2>, tc::GemmSmem<NT>::BYTES);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)w.num_items);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = tc::GemmSmem<NT>::BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, tc::tc_gemm_kernel<NT, true, 2>, maps, w, dx_part, batch);
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();            // a device that cannot place CTA pairs (partitioned GPU): same kernel without the cluster
    return launch_single();
}

static cudaError_t launch_skin(int mode, const TcConstMaps& cm, const float* ae_hi, const float* ae_lo, const float* in, float* out0,
                               float* out1, float* out2, int batch, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    tc::SkinMaps maps;
    maps.w_hi = cm.w_hi;
    maps.w_lo = cm.w_lo;
    if (!make_map(&maps.ae_hi, ae_hi, 32, (uint64_t)batch * 12, 0, tc::SK_N) || !make_map(&maps.ae_lo, ae_lo, 32, (uint64_t)batch * 12, 0, tc::SK_N))
        return cudaErrorInvalidValue;
    const int nchunks = (batch + tc::SK_NB - 1) / tc::SK_NB;
    // One CTA per SM at a time (181 KB of shared memory), every CTA does nchunks / groups chunks: the grid of 54 x groups CTAs
    // should fill whole waves.  Among 2..12 sample groups take the one that wastes the least of its last wave (148 SMs:
    // 8 groups = 432 CTAs = 2.92 waves; the former "two CTAs per SM" choice of 6 groups = 2.19 waves idled 27 % of the SMs).
    const int sms = sm_count();
    int groups = 1;
    double best = 0.0;
    for (int g = 2; g <= 12 && g <= nchunks; ++g) {
        const int ctas = tc::SK_VBLOCKS * g, waves = (ctas + sms - 1) / sms;
        const double eff = (double)ctas / ((double)waves * sms);
        if (eff > best + 1e-9) { best = eff; groups = g; }
    }
    cudaError_t e = mode == 0 ? opt_in(tc::tc_skin_kernel<0>, tc::SK_SMEM) : opt_in(tc::tc_skin_kernel<1>, tc::SK_SMEM);
    if (e != cudaSuccess) return e;
    const int grid = tc::SK_VBLOCKS * groups;
    if (mode == 0) tc::tc_skin_kernel<0><<<grid, 384, tc::SK_SMEM, stream>>>(maps, in, out0, out1, out2, batch, groups);
    else tc::tc_skin_kernel<1><<<grid, 384, tc::SK_SMEM, stream>>>(maps, in, out0, out1, out2, batch, groups);
    return cudaGetLastError();
}

cudaError_t launch_skin_forward(const TcConstMaps& cm, const float* ae_hi, const float* ae_lo, const float* vposed, float* verts,
                                int batch, cudaStream_t stream) {
    return launch_skin(0, cm, ae_hi, ae_lo, vposed, verts, nullptr, nullptr, batch, stream);
}
cudaError_t launch_skin_backward(const TcConstMaps& cm, const float* ae_hi, const float* ae_lo, const float* dverts, float* dvp,
                                 float* dverts_padded, int batch, cudaStream_t stream) {
    return launch_skin(1, cm, ae_hi, ae_lo, dverts, dvp, nullptr, dverts_padded, batch, stream);
}

int tc_dA_splits(int batch) {
    const int mtiles = (batch + tc::BM - 1) / tc::BM;
    int n = (sm_count() + mtiles - 1) / mtiles;
    if (n < 16) n = 16;                   // <= 14 k-blocks (168 MMAs) per accumulation chain
    if (n > kMaxSplitA) n = kMaxSplitA;
    return n;
}

cudaError_t launch_dA(const TcConstMaps& cm, const float* dverts_padded, const float* vposed, float* dA_part, int batch, int nsplit,
                      cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    tc::DaMaps maps;
    maps.wT_hi = cm.wT_hi;
    maps.wT_lo = cm.wT_lo;
    if (!make_map(&maps.dv, dverts_padded, kVpPitch, (uint64_t)batch, 0, tc::BM) || !make_map(&maps.vp, vposed, kVpPitch, (uint64_t)batch, 0, tc::BM))
        return cudaErrorInvalidValue;
    cudaError_t e = opt_in(tc::tc_dA_kernel, tc::DA_SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid(nsplit, (batch + tc::BM - 1) / tc::BM);
    tc::tc_dA_kernel<<<grid, 384, tc::DA_SMEM, stream>>>(maps, dA_part, batch, nsplit);
    return cudaGetLastError();
}

}  // namespace smplb200
