// GEMM engine of the pair fit kernel (fit_pair.cuh): the three per-iteration contractions of the SMPLify loop
//   prior  Pd[576 x 32]  = Psym[576 x 72]  . bp[72 x 32]        (smplify/prior.py:181-196, all 8 mixture components)
//   fwd    Q [704 x 32]  = Cf^T[704 x 224] . x [224 x 32]       (the folded joint model)
//   bwd    dx[224 x 32]  = Cf [224 x 704]  . dQ[704 x 32]
// on the tcgen05 tensor cores of a CTA PAIR (cta_group::2, M = 256): the two CTAs of a 2-CTA cluster fit 16 samples each;
// the 32 samples are the N dimension of one MMA, each CTA supplying its own 16 rows of the B operand from its shared memory.
// The fp32 constants are the A operand: each CTA streams ITS 128-row half of every 256-row tile from L2 with ld.global.nc,
// splits it in registers into the 19 bits the tensor core reads (hi) and the remainder (lo) and writes both into a ring in
// TMEM (TS-mode MMA, A from TMEM), so every constant leaves L2 once per PAIR and iteration - half the L2 -> SM traffic of
// one CTA per 16 samples, which is what bounds the loop once the arithmetic is on the tensor pipe.  3xTF32 with the hi.hi
// products and the lo.hi + hi.lo corrections in SEPARATE fp32 accumulators (the tensor pipe truncates its accumulator after
// every MMA; the short hi.hi chain keeps that below 1e-6 relative) and, for the K = 704 contraction, four partial
// accumulator pairs over K ranges that the epilogue adds in fp32.
//
// Warp roles during a call (448 threads): warps 0-7 generate the A operand (two groups of four warps alternate chunks of 32
// k-values), warp 12 of the leader CTA issues the MMAs for both CTAs, warps 8-11 drain the accumulators (their lane
// quarter each) into their own and - through distributed shared memory - the peer's tile state, warp 13 is free for the
// kinematic chain sweeps.  All waits are bounded and trap.
#pragma once
#include "smpl_common.h"
#include "tc_common.cuh"

namespace smplb200 {
namespace pg {

using namespace tc;

constexpr int kThreads = 448;
constexpr int kEpiWarp0 = 8, kMmaWarp = 12, kChainWarp = 13;
constexpr int CH = 32;                          // k values per chunk = one SWIZZLE_128B atom of the B operand
constexpr int NBUF = 3;                         // TMEM ring slots of 2 * CH columns (hi | lo)
constexpr int kAccCol = NBUF * 2 * CH;          // accumulators start here (column 192)
constexpr int kTmemCols = 512;
constexpr int NS = 32;                          // MMA N: 2 CTAs x 16 sample rows
constexpr int kAtomFloats = 16 * 32;            // one B atom: 16 sample rows x 32 k (2 KB)

// ---- problem shapes ----------------------------------------------------------------------------------------------------
constexpr int kPriorTiles = 3, kPriorChunks = 3, kPriorLastSteps = 1;      // 576 rows -> 3 x 256; K = 72 = 2 chunks + 1 k-step
constexpr int kFwdTiles = 3, kFwdChunks = kXPad / CH;                      // 704 rows -> 3 x 256; K = 224 = 7 chunks
constexpr int kBwdChunks = kQPad / CH;                                     // 224 rows -> 1 x 256; K = 704 = 22 chunks
constexpr int kBwdParts = 4;                                               // partial accumulator pairs over K ranges
constexpr int kFwdCallChunks = kPriorTiles * kPriorChunks + kFwdTiles * kFwdChunks;      // 30
__host__ __device__ constexpr int bwd_part_begin(int p) { return p == 0 ? 0 : (p == 1 ? 6 : (p == 2 ? 12 : (p == 3 ? 17 : 22))); }
// packed constant arrays: [tile][half][k chunk][8 float4 per row][128 rows] float4
constexpr size_t kPackedPriorFloats = (size_t)kPriorTiles * 2 * kPriorChunks * CH * 128;
constexpr size_t kPackedFwdFloats = (size_t)kFwdTiles * 2 * kFwdChunks * CH * 128;
constexpr size_t kPackedBwdFloats = (size_t)1 * 2 * kBwdChunks * CH * 128;

struct Consts {                  // device pointers (model blob)
    const float4* prior;         // rows R = 256 t + 128 h + r -> (g, i) = (R / 72, R % 72), k = j
    const float4* fwd;           // rows R -> n, k = m          (Cf[m][n])
    const float4* bwd;           // rows R = 128 h + r -> m, k = n
};

struct Bars {                    // shared memory, identical offsets in both CTAs of the pair
    uint64_t full[NBUF];         // LEADER's copy is used: 8 arrivals (4 generator warps of each CTA)
    uint64_t empty[NBUF];        // 1 arrival: multicast commit of the MMAs that read the slot
    uint64_t acc_full[6];        // 1 arrival: multicast commit after the last MMA into an accumulator (set)
    uint64_t epi;                // 8 arrivals: the epilogue warps of both CTAs
    uint32_t tmem_base;
    uint32_t pad;
};

// Running counters of barrier uses: identical in every thread of both CTAs (every call has a fixed shape), so that phase
// parities never need to be communicated.
struct State {
    uint32_t chunks;             // ring chunks issued by the calls before this one
    uint32_t fwd_calls, bwd_calls, epi_syncs;
};

// ---- cluster / DSMEM primitives ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local_cluster(uint32_t bar) {      // cluster-scope release on a local barrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait with cluster-scope acquire (the producers may sit in the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    printf("pg: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void st_remote_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_remote_v4(uint32_t cluster_addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---- tcgen05, cta_group::2 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kTmemCols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem of both CTAs], M = 256 (128 rows per CTA), tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma2_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}

// ---- B operand layout -----------------------------------------------------------------------------------------------------
// float offset of (sample row s < 16, k) inside a K-major SWIZZLE_128B operand made of atoms of 32 k (tc_common.cuh swz128)
__host__ __device__ __forceinline__ int b_off(int s, int k) {
    return (k >> 5) * kAtomFloats + s * 32 + ((((k & 31) >> 2) ^ (s & 7)) << 2) + (k & 3);
}
// the 19 bits of an fp32 the tensor core reads as tf32, and the remainder (exact)
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

__device__ __forceinline__ void bars_init(Bars* b) {
    for (int i = 0; i < NBUF; ++i) { mbar_init(smem_u32(&b->full[i]), 8); mbar_init(smem_u32(&b->empty[i]), 1); }
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&b->acc_full[i]), 1);
    mbar_init(smem_u32(&b->epi), 8);
    fence_barrier_init();
}

// All eight epilogue warps of the pair (4 + 4) meet: DSMEM writes issued before it by either CTA's epilogue warps are visible
// to both afterwards.  One elected lane per warp arrives on both CTAs' barriers.
__device__ __forceinline__ void epi_pair_sync(Bars* b, State& st, uint32_t rank) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t bar = smem_u32(&b->epi);
        mbar_arrive_local_cluster(bar);
        mbar_arrive_remote(mapa(bar, rank ^ 1u));
    }
    mbar_wait_cluster(smem_u32(&b->epi), st.epi_syncs & 1u);
    ++st.epi_syncs;
}

// ---- generator warps ---------------------------------------------------------------------------------------------------------
// One chunk = 32 k values of this thread's row: 8 float4.
struct Chunk { float4 v[8]; };
__device__ __forceinline__ void chunk_load(Chunk& c, const float4* p) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c.v[i] = __ldg(p + i * 128);
}
// split, wait for the ring slot, write hi | lo into TMEM, hand the slot to the MMA warp of the leader CTA
__device__ __forceinline__ void chunk_publish(const Chunk& c, Bars* b, uint32_t g, uint32_t lane_addr, uint32_t full_leader_base) {
    const uint32_t slot = g % NBUF, use = g / NBUF;
    float hi[CH], lo[CH];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float v[4] = {c.v[i].x, c.v[i].y, c.v[i].z, c.v[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            hi[4 * i + e] = tf32_trunc(v[e]);
            lo[4 * i + e] = v[e] - hi[4 * i + e];
        }
    }
    if (use > 0) mbar_wait_cluster(smem_u32(&b->empty[slot]), (use - 1) & 1u);
    tc_fence_after();
    const uint32_t ta = lane_addr + slot * 2 * CH;
    tmem_st16(ta, hi); tmem_st16(ta + 16, hi + 16);
    tmem_st16(ta + CH, lo); tmem_st16(ta + CH + 16, lo + 16);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_remote(full_leader_base + slot * 8);
}

// packed-array address of chunk c of the forward call (prior tiles first) / of the backward call, for this thread's row
__device__ __forceinline__ const float4* fwd_chunk_ptr(const Consts& K, int c, uint32_t rank, int row) {
    if (c < kPriorTiles * kPriorChunks) {
        const int t = c / kPriorChunks, kc = c % kPriorChunks;
        return K.prior + ((size_t)((t * 2 + rank) * kPriorChunks + kc) * 8) * 128 + row;
    }
    c -= kPriorTiles * kPriorChunks;
    const int t = c / kFwdChunks, kc = c % kFwdChunks;
    return K.fwd + ((size_t)((t * 2 + rank) * kFwdChunks + kc) * 8) * 128 + row;
}
__device__ __forceinline__ const float4* bwd_chunk_ptr(const Consts& K, int c, uint32_t rank, int row) {
    return K.bwd + ((size_t)(rank * kBwdChunks + c) * 8) * 128 + row;
}

template <bool FWD>
__device__ __forceinline__ void generator_run(const Consts& K, Bars* b, const State& st, uint32_t rank, uint32_t tmem_base) {
    constexpr int NC = FWD ? kFwdCallChunks : kBwdChunks;          // even
    const int warp = threadIdx.x >> 5, wg = warp >> 2, row = (threadIdx.x & 127);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t full_leader = mapa(smem_u32(&b->full[0]), 0);
    auto ptr = [&](int c) { return FWD ? fwd_chunk_ptr(K, c, rank, row) : bwd_chunk_ptr(K, c, rank, row); };
    Chunk c0, c1;
    chunk_load(c0, ptr(wg));
    chunk_load(c1, ptr(wg + 2));
    constexpr int MINE = NC / 2;
#pragma unroll 1
    for (int i = 0; i < MINE; i += 2) {
        chunk_publish(c0, b, st.chunks + wg + 2 * i, lane_addr, full_leader);
        if (i + 2 < MINE) chunk_load(c0, ptr(wg + 2 * (i + 2)));
        if (i + 1 < MINE) {
            chunk_publish(c1, b, st.chunks + wg + 2 * (i + 1), lane_addr, full_leader);
            if (i + 3 < MINE) chunk_load(c1, ptr(wg + 2 * (i + 3)));
        }
    }
}

// ---- MMA warp (leader CTA) ---------------------------------------------------------------------------------------------------
// accumulator columns (from tmem_base): forward call: prior tile t at kAccCol + 32 t (hi.hi and corrections together),
// forward tile t main at kAccCol + 96 + 64 t, corrections 32 further; backward call: part p main at kAccCol + 64 p.
__device__ __forceinline__ void issue_chunk(uint32_t tmem_base, uint32_t slot, uint32_t d_main, uint32_t d_corr, uint64_t bh, uint64_t bl,
                                            int ksteps, bool first_of_acc, bool shared_acc, uint32_t idesc) {
    const uint32_t a_hi = tmem_base + slot * 2 * CH, a_lo = a_hi + CH;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        if (ks < ksteps) {
            const bool first = first_of_acc && ks == 0;
            umma2_tf32_ts(d_main, a_hi + 8 * ks, bh + (uint64_t)(2 * ks), idesc, first ? 0u : 1u);
            umma2_tf32_ts(d_corr, a_lo + 8 * ks, bh + (uint64_t)(2 * ks), idesc, (first && !shared_acc) ? 0u : 1u);
            umma2_tf32_ts(d_corr, a_hi + 8 * ks, bl + (uint64_t)(2 * ks), idesc, 1u);
        }
    }
}

struct BOperands { uint32_t prior_hi, prior_lo, x_hi, x_lo, dq_hi, dq_lo; };      // smem byte addresses (1024-byte aligned atoms)

template <bool FWD>
__device__ __forceinline__ void mma_run(Bars* b, const State& st, const BOperands& B, uint32_t tmem_base) {
    constexpr uint32_t idesc = instr_desc(256, NS);
    const uint32_t call = FWD ? st.fwd_calls : st.bwd_calls;
    (void)call;
    if (FWD) {
#pragma unroll 1
        for (int c = 0; c < kFwdCallChunks; ++c) {
            const uint32_t g = st.chunks + c, slot = g % NBUF, use = g / NBUF;
            const bool prior = c < kPriorTiles * kPriorChunks;
            const int cc = prior ? c : c - kPriorTiles * kPriorChunks;
            const int t = prior ? cc / kPriorChunks : cc / kFwdChunks, kc = prior ? cc % kPriorChunks : cc % kFwdChunks;
            const int nchunks = prior ? kPriorChunks : kFwdChunks;
            const int ksteps = (prior && kc == kPriorChunks - 1) ? kPriorLastSteps : 4;
            const uint32_t d_main = tmem_base + kAccCol + (prior ? 32 * t : 96 + 64 * t), d_corr = prior ? d_main : d_main + 32;
            const uint64_t bh = smem_desc((prior ? B.prior_hi : B.x_hi) + kc * kAtomFloats * 4),
                           bl = smem_desc((prior ? B.prior_lo : B.x_lo) + kc * kAtomFloats * 4);
            mbar_wait_cluster(smem_u32(&b->full[slot]), use & 1u);
            tc_fence_after();
            if (elect_one()) {
                issue_chunk(tmem_base, slot, d_main, d_corr, bh, bl, ksteps, kc == 0, prior, idesc);
                umma2_commit(smem_u32(&b->empty[slot]));
                if (kc == nchunks - 1) umma2_commit(smem_u32(&b->acc_full[prior ? t : 3 + t]));
            }
            __syncwarp();
        }
    } else {
#pragma unroll 1
        for (int c = 0; c < kBwdChunks; ++c) {
            const uint32_t g = st.chunks + c, slot = g % NBUF, use = g / NBUF;
            const int p = c < bwd_part_begin(1) ? 0 : (c < bwd_part_begin(2) ? 1 : (c < bwd_part_begin(3) ? 2 : 3));
            const uint32_t d_main = tmem_base + kAccCol + 64 * p, d_corr = d_main + 32;
            const uint64_t bh = smem_desc(B.dq_hi + c * kAtomFloats * 4), bl = smem_desc(B.dq_lo + c * kAtomFloats * 4);
            mbar_wait_cluster(smem_u32(&b->full[slot]), use & 1u);
            tc_fence_after();
            if (elect_one()) {
                issue_chunk(tmem_base, slot, d_main, d_corr, bh, bl, 4, c == bwd_part_begin(p), false, idesc);
                umma2_commit(smem_u32(&b->empty[slot]));
                if (c == bwd_part_begin(p + 1) - 1) umma2_commit(smem_u32(&b->acc_full[p]));
            }
            __syncwarp();
        }
    }
}

// ---- epilogue helpers ------------------------------------------------------------------------------------------------------
// this thread's 32 accumulator columns (16 samples of CTA 0, then 16 of CTA 1) of one accumulator
__device__ __forceinline__ void acc_load32(uint32_t taddr, float* v) {
    tmem_ld32(taddr, v);
    tmem_ld_wait();
}
__device__ __forceinline__ void acc_wait(Bars* b, int idx, uint32_t use) {
    mbar_wait_cluster(smem_u32(&b->acc_full[idx]), use & 1u);
    tc_fence_after();
}

__device__ __forceinline__ void state_advance_fwd(State& st) { st.chunks += kFwdCallChunks; ++st.fwd_calls; }
__device__ __forceinline__ void state_advance_bwd(State& st) { st.chunks += kBwdChunks; ++st.bwd_calls; }
// acc_full[i] is used once per forward call (i < 6) and once per backward call (i < 4)
__device__ __forceinline__ uint32_t acc_use_fwd(const State& st, int i) { return i < kBwdParts ? st.fwd_calls + st.bwd_calls : st.fwd_calls; }
__device__ __forceinline__ uint32_t acc_use_bwd(const State& st, int i) { (void)i; return st.fwd_calls + st.bwd_calls; }

}  // namespace pg
}  // namespace smplb200
