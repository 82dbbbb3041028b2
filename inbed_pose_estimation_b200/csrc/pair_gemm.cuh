// GEMM engine of the pair fit kernel (fit_pair.cuh): the three per-iteration contractions of the SMPLify loop
//   prior  Pd[576 x 32]  = Psym[576 x 72]  . bp[72 x 32]        (smplify/prior.py:181-196, all 8 mixture components)
//   fwd    Q [704 x 32]  = Cf^T[704 x 224] . x [224 x 32]       (the folded joint model)
//   bwd    dx[224 x 32]  = Cf [224 x 704]  . dQ[704 x 32]
// on the tcgen05 tensor cores of a CTA PAIR (cta_group::2, M = 256): the two CTAs of a 2-CTA cluster fit 16 samples each;
// the 32 samples are the N dimension of one MMA, each CTA supplying its own 16 rows of the B operand from its shared memory.
// The fp32 constants are the A operand: each CTA streams ITS 128-row half of every 256-row tile from L2 with ld.global.nc,
// splits it in registers into the 19 bits the tensor core reads (hi) and the remainder (lo) and writes both into a ring in
// TMEM (TS-mode MMA, A from TMEM), so every constant leaves L2 once per PAIR and iteration.  3xTF32 with the hi.hi products
// and the lo.hi + hi.lo corrections in SEPARATE fp32 accumulators (the tensor pipe truncates its accumulator after every
// MMA; the short hi.hi chain keeps that below 1e-6 relative) and, for the K = 704 contraction, three partial accumulator
// pairs over K ranges that the epilogue adds in fp32.
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
constexpr int NBUF = 4;                         // TMEM ring slots of 2 * CH columns (hi | lo).  EVEN on purpose: chunk g goes to
                                                // slot g % NBUF and is produced by warp group g % 2, so a slot always belongs to
                                                // the same group and its warps see every phase of the slot's `empty` barrier (a
                                                // parity wait must never fall two phases behind).  Measured alternatives: 2 slots
                                                // of 64 k (half the hand-shakes) are slower - the ring is then too shallow for
                                                // the slot round trip; 5 slots need per-slot ownership to be given up
constexpr int kAccCol = NBUF * 2 * CH;          // accumulators start here (column 256); 192 columns used
constexpr int kTmemCols = 512;
constexpr int NS = 32;                          // MMA N: 2 CTAs x 16 sample rows
constexpr int kAtomFloats = 16 * 32;            // one B atom: 16 sample rows x 32 k (2 KB)

// ---- problem shapes ----------------------------------------------------------------------------------------------------
static_assert(CH == kPgChunk, "chunk size of the packed constants");
static_assert(NBUF % 2 == 0 && kAccCol + 192 <= kTmemCols, "ring / accumulator split of TMEM");
constexpr int kPriorTiles = kPgPriorTiles, kPriorChunks = kPgPriorChunks, kPriorLastSteps = 1;      // 576 rows -> 3 x 256; K = 72 = 2 chunks + 1 k-step
constexpr int kFwdTiles = kPgFwdTiles, kFwdChunks = kPgFwdChunks;          // 704 rows -> 3 x 256; K = 224 = 7 chunks
constexpr int kBwdChunks = kPgBwdChunks;                                   // 224 rows -> 1 x 256; K = 704 = 22 chunks
constexpr int kBwdParts = 3;                                               // partial accumulator pairs over K ranges
constexpr int kFwdCallChunks = kPriorTiles * kPriorChunks + kFwdTiles * kFwdChunks;      // 30
__host__ __device__ constexpr int bwd_part_begin(int p) { return p == 0 ? 0 : (p == 1 ? 8 : (p == 2 ? 15 : 22)); }

struct Consts {                  // device pointers (model blob)
    const float4* prior;         // rows R = 256 t + 128 h + r -> (g, i) = (R / 72, R % 72), k = j
    const float4* fwd;           // rows R -> n, k = m          (Cf[m][n])
    const float4* bwd;           // rows R = 128 h + r -> m, k = n
};

struct Bars {                    // shared memory, identical offsets in both CTAs of the pair; re-initialised before every call
    uint64_t full[NBUF];         // LEADER's copy is used: 8 arrivals (4 generator warps of each CTA)
    uint64_t empty[NBUF];        // 1 arrival: multicast commit of the MMAs that read the slot
    uint64_t acc_full[6];        // 1 arrival: multicast commit after the last MMA into an accumulator (set)
    uint64_t acc_free[2];        // LEADER's copy: 8 arrivals (epilogue warps of both CTAs): prior tiles 0+1 / 2 have left TMEM
    uint64_t epi;                // 8 arrivals: the epilogue warps of both CTAs
    uint32_t tmem_base;
    uint32_t pad;
};

// ---- cluster / DSMEM primitives ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrive on a barrier in (possibly) another CTA of the cluster.  Default semantics (release at CTA scope): what it publishes here
// is TMEM content, ordered by the tcgen05 fences; a cluster-scope release would drain the generator's outstanding prefetches.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// the same with a cluster-scope release: for hand-offs of ordinary (distributed) shared-memory data
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local_cluster(uint32_t bar) {      // cluster-scope release on a local barrier
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#if !defined(PG_SPIN_LOG2)
#define PG_SPIN_LOG2 22
#endif
static __device__ int g_pg_call_index;          // diagnostics only: number of engine calls block 0 has started
// bounded wait (CTA-scope acquire): the engine's own copy, so that debugging builds can shorten it for this kernel only
__device__ __forceinline__ void pg_wait(uint32_t bar, uint32_t parity, int what) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << PG_SPIN_LOG2); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    if ((threadIdx.x & 31) == 0)
        printf("pg: wait %d timed out (block %d warp %d bar %u parity %u, engine call %d of block 0)\n", what, blockIdx.x, threadIdx.x >> 5, bar,
               parity, g_pg_call_index);
    __trap();
}
// non-blocking probe of a phase: issued ahead of time so that its latency hides behind other work
__device__ __forceinline__ uint32_t pg_probe(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
// bounded wait with cluster-scope acquire (the producers of ordinary shared-memory data may sit in the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << PG_SPIN_LOG2); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    if ((threadIdx.x & 31) == 0)
        printf("pg: pair-sync wait timed out (block %d warp %d bar %u parity %u)\n", blockIdx.x, threadIdx.x >> 5, bar, parity);
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
// arrive on the mbarrier at this offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
// low 32 bits of the K-major SWIZZLE_128B descriptor of an operand at shared-memory byte address `addr` (tc_common.cuh
// smem_desc); advancing by one 32-byte k-step adds 2, by one atom of 16 rows adds 128.  The high word is a constant.
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);      // SBO = 1024 bytes, version 1, SWIZZLE_128B
static_assert(((uint64_t)kDescHi << 32) == ((64ull << 32) | (1ull << 46) | (2ull << 61)), "descriptor high word");
// D[tmem] (+)= A[tmem] . B[smem of both CTAs], M = 256 (128 rows per CTA), tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}

// ---- B operand layout -----------------------------------------------------------------------------------------------------
// float offset of (sample row s < 16, k) inside a K-major SWIZZLE_128B operand made of atoms of 32 k (tc_common.cuh swz128)
__host__ __device__ __forceinline__ int b_off(int s, int k) {
    return (k >> 5) * kAtomFloats + s * 32 + ((((k & 31) >> 2) ^ (s & 7)) << 2) + (k & 3);
}
// the 19 bits of an fp32 the tensor core reads as tf32, and the remainder (exact)
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// Every call starts from freshly initialised barriers (one thread per CTA, before the cluster barrier that precedes the call):
// slot numbers and phase parities inside a call are then plain functions of the chunk index.
__device__ __forceinline__ void mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
// `again`: the objects exist already (an mbarrier must be invalidated before it is initialised a second time)
__device__ __forceinline__ void bars_init(Bars* b, bool again) {
    if (again) {
        for (int i = 0; i < NBUF; ++i) { mbar_inval(smem_u32(&b->full[i])); mbar_inval(smem_u32(&b->empty[i])); }
        for (int i = 0; i < 6; ++i) mbar_inval(smem_u32(&b->acc_full[i]));
        for (int i = 0; i < 2; ++i) mbar_inval(smem_u32(&b->acc_free[i]));
        mbar_inval(smem_u32(&b->epi));
    }
    for (int i = 0; i < NBUF; ++i) { mbar_init(smem_u32(&b->full[i]), 8); mbar_init(smem_u32(&b->empty[i]), 1); }
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&b->acc_full[i]), 1);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&b->acc_free[i]), 8);
    mbar_init(smem_u32(&b->epi), 8);
    fence_barrier_init();
}

// All eight epilogue warps of the pair (4 + 4) meet (once per call): DSMEM writes issued before it by either CTA's epilogue
// warps are visible to both afterwards.  One elected lane per warp arrives on both CTAs' barriers.
__device__ __forceinline__ void epi_pair_sync(Bars* b, uint32_t rank) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t bar = smem_u32(&b->epi);
        mbar_arrive_local_cluster(bar);
        mbar_arrive_remote_release(mapa(bar, rank ^ 1u));
    }
    mbar_wait_cluster(smem_u32(&b->epi), 0u);
}

#if defined(SMPLB200_PHASE_CLOCKS)
#define PG_WAIT_BEGIN() const long long pgw_t0 = clock64()
#define PG_WAIT_END(acc) (acc) += clock64() - pgw_t0
#else
#define PG_WAIT_BEGIN() ((void)0)
#define PG_WAIT_END(acc) ((void)0)
#endif
// Event trace of one forward call of CTA 0 (profiling builds with -DPG_TRACE): (event id, chunk) -> clock64 of SM 0
#if defined(PG_TRACE)
static __device__ long long g_pg_trace[8 * 64];
static __device__ int g_pg_trace_on;
// every warp that records events caches the switch in a register once per call (PG_TRACE_CALL_BEGIN): no loads in the loops
#define PG_TRACE_CALL_BEGIN() const bool pg_trace_here = ::smplb200::pg::g_pg_trace_on && blockIdx.x == 0 && (threadIdx.x & 31) == 0
#define PG_TRACE_EVT(ev, chunk) do { if (pg_trace_here && (chunk) < 64) ::smplb200::pg::g_pg_trace[(ev) * 64 + (chunk)] = clock64(); } while (0)
#else
#define PG_TRACE_CALL_BEGIN() ((void)0)
#define PG_TRACE_EVT(ev, chunk) ((void)0)
#endif

// ---- generator warps ---------------------------------------------------------------------------------------------------------
// One chunk = 32 k values of this thread's row: 8 float4.
struct Chunk { float4 v[8]; };
__device__ __forceinline__ void chunk_load(Chunk& c, const float4* p) {
#if defined(PG_NO_LOAD)                                      // timing discriminator (wrong results): constants not fetched
    (void)p;
#pragma unroll
    for (int i = 0; i < 8; ++i) c.v[i] = make_float4(0.25f, -0.5f, 0.125f, 1.f);
#else
#pragma unroll
    for (int i = 0; i < 8; ++i) c.v[i] = __ldg(p + i * 128);
#endif
}
// split, wait for the ring slot, write hi | lo into TMEM, hand the slot to the MMA warp of the leader CTA; g = chunk index
// inside the call
__device__ __forceinline__ void chunk_publish(const Chunk& c, Bars* b, uint32_t g, uint32_t lane_addr, uint32_t full_leader_base,
                                              long long& waited, long long& stored, int cidx, bool pg_trace_here) {
    (void)pg_trace_here;
    PG_TRACE_EVT(0, cidx);
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
    {
        PG_WAIT_BEGIN();
        if (use > 0) pg_wait(smem_u32(&b->empty[slot]), (use - 1) & 1u, 100 + (int)g);
        PG_WAIT_END(waited);
    }
    PG_TRACE_EVT(1, cidx);
    PG_WAIT_BEGIN();
    tc_fence_after();
    const uint32_t ta = lane_addr + slot * 2 * CH;
    tmem_st16(ta, hi); tmem_st16(ta + 16, hi + 16);
    tmem_st16(ta + CH, lo); tmem_st16(ta + CH + 16, lo + 16);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive_remote(full_leader_base + slot * 8);
    PG_WAIT_END(stored);
    PG_TRACE_EVT(2, cidx);
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
__device__ __forceinline__ void generator_run(const Consts& K, Bars* b, uint32_t rank, uint32_t tmem_base) {
    constexpr int NC = FWD ? kFwdCallChunks : kBwdChunks;          // even
    const int warp = threadIdx.x >> 5, wg = warp >> 2, row = (threadIdx.x & 127);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t full_leader = mapa(smem_u32(&b->full[0]), 0);
    auto ptr = [&](int c) { return FWD ? fwd_chunk_ptr(K, c, rank, row) : bwd_chunk_ptr(K, c, rank, row); };
#if defined(PG_TRACE)
    PG_TRACE_CALL_BEGIN();
    const bool pg_trace_gen = pg_trace_here && FWD;
#else
    const bool pg_trace_gen = false;
#endif
    Chunk c0, c1;
    chunk_load(c0, ptr(wg));
    chunk_load(c1, ptr(wg + 2));
    constexpr int MINE = NC / 2;
    long long waited = 0, stored = 0;
#pragma unroll 1
    for (int i = 0; i < MINE; i += 2) {
        chunk_publish(c0, b, wg + 2 * i, lane_addr, full_leader, waited, stored, (warp & 3) == 0 ? wg + 2 * i : 64, pg_trace_gen);
        if (i + 2 < MINE) chunk_load(c0, ptr(wg + 2 * (i + 2)));
        if (i + 1 < MINE) {
            chunk_publish(c1, b, wg + 2 * (i + 1), lane_addr, full_leader, waited, stored, (warp & 3) == 0 ? wg + 2 * (i + 1) : 64, pg_trace_gen);
            if (i + 3 < MINE) chunk_load(c1, ptr(wg + 2 * (i + 3)));
        }
    }
    // drain: the commits of the last use of this group's slots are waited for too, so that no arrival of this call can still
    // be on its way when the barriers are re-initialised for the next one
#pragma unroll
    for (int k = 0; k < NBUF / 2; ++k) {                                 // this group's slots: wg, wg + 2
        const int slot = wg + 2 * k, last_use = (NC - 1 - slot) / NBUF;
        pg_wait(smem_u32(&b->empty[slot]), (uint32_t)last_use & 1u, 150 + slot);
    }
#if defined(SMPLB200_PHASE_CLOCKS)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&g_phase_clocks[FWD ? 24 : 28], (unsigned long long)waited);
        atomicAdd(&g_phase_clocks[FWD ? 25 : 29], (unsigned long long)stored);
    }
#endif
}

// ---- MMA warp (leader CTA) ---------------------------------------------------------------------------------------------------
// accumulator columns (from tmem_base + kAccCol, 192 columns): forward call: prior tile t at 32 t (hi.hi and corrections
// together), forward tile t main at 64 t, corrections 32 further - forward tiles 0 / 1 reuse the columns of the prior tiles
// 0+1 / 2, which the epilogue warps of both CTAs release through acc_free[0 / 1]; backward call: part p main at 64 p.
//
// One warp issues everything, and a lone warp pays the full latency of every dependent instruction: the per-chunk bookkeeping
// is a rotating slot counter and additive descriptor updates, nothing else.
struct BOperands { uint32_t prior_hi, prior_lo, x_hi, x_lo, dq_hi, dq_lo; };      // smem byte addresses (1024-byte aligned atoms)

struct RingPos {
    uint32_t slot = 0, parity = 0;
    uint32_t ready = 0;           // result of the probe of (slot, parity) issued while the previous chunk's MMAs were being issued
    __device__ __forceinline__ void advance() { if (++slot == (uint32_t)NBUF) { slot = 0; parity ^= 1u; } }
};

// one chunk: wait for it, 3 MMAs per k-step (hi.hi -> d_main; lo.hi and hi.lo -> d_corr), release the slot
template <int KSTEPS>
__device__ __forceinline__ void mma_chunk(Bars* b, RingPos& rp, uint32_t tmem_base, uint32_t d_main, uint32_t d_corr, uint32_t bh, uint32_t bl,
                                          uint32_t acc_first_main, uint32_t acc_first_corr, uint32_t idesc, long long& waited, int cidx,
                                          bool pg_trace_here) {
    (void)pg_trace_here;
    PG_TRACE_EVT(3, cidx);
    if (!rp.ready) {
        PG_WAIT_BEGIN();
        pg_wait(smem_u32(&b->full[rp.slot]), rp.parity, 200 + cidx);
        PG_WAIT_END(waited);
    }
    PG_TRACE_EVT(4, cidx);
    tc_fence_after();
    const uint32_t a_hi = tmem_base + rp.slot * 2 * CH, a_lo = a_hi + CH, empty_bar = smem_u32(&b->empty[rp.slot]);
    rp.advance();
    rp.ready = pg_probe(smem_u32(&b->full[rp.slot]), rp.parity);      // the next chunk: the answer arrives while this one is issued
    if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            umma2(d_main, a_hi + 8 * ks, bh + 2 * ks, idesc, ks ? 1u : acc_first_main);
#if !defined(PG_ONE_MMA)                                     // timing discriminator (wrong results): hi.hi products only
            umma2(d_corr, a_lo + 8 * ks, bh + 2 * ks, idesc, ks ? 1u : acc_first_corr);
            umma2(d_corr, a_hi + 8 * ks, bl + 2 * ks, idesc, 1u);
#endif
        }
#if !defined(PG_NO_EMPTY_COMMIT)                              // timing discriminator (wrong results; generators run free)
        umma2_commit(empty_bar);
#endif
    }
    __syncwarp();
    PG_TRACE_EVT(5, cidx);
}

__device__ __forceinline__ void mma_commit_acc(Bars* b, int idx) {
    if (elect_one()) umma2_commit(smem_u32(&b->acc_full[idx]));
    __syncwarp();
}

template <bool FWD>
__device__ __forceinline__ void mma_run(Bars* b, const BOperands& B, uint32_t tmem_base) {
    constexpr uint32_t idesc = instr_desc(256, NS);
    long long waited = 0;
    RingPos rp;
#if defined(PG_TRACE)
    PG_TRACE_CALL_BEGIN();
    const bool pg_trace_mma = pg_trace_here && FWD;
#else
    const bool pg_trace_mma = false;
#endif
    const uint32_t acc = tmem_base + kAccCol;
    if (FWD) {
        const uint32_t ph = desc_lo(B.prior_hi), pl = desc_lo(B.prior_lo), xh = desc_lo(B.x_hi), xl = desc_lo(B.x_lo);
#pragma unroll 1
        for (int t = 0; t < kPriorTiles; ++t) {                  // K = 72: two whole chunks and one k-step; one accumulator
            const uint32_t d = acc + 32 * t;
            mma_chunk<4>(b, rp, tmem_base, d, d, ph, pl, 0u, 1u, idesc, waited, 3 * t, pg_trace_mma);
            mma_chunk<4>(b, rp, tmem_base, d, d, ph + 128, pl + 128, 1u, 1u, idesc, waited, 3 * t + 1, pg_trace_mma);
            mma_chunk<kPriorLastSteps>(b, rp, tmem_base, d, d, ph + 256, pl + 256, 1u, 1u, idesc, waited, 3 * t + 2, pg_trace_mma);
            mma_commit_acc(b, t);
        }
#pragma unroll 1
        for (int t = 0; t < kFwdTiles; ++t) {
            const uint32_t d = acc + 64 * t;
            if (t < 2) pg_wait(smem_u32(&b->acc_free[t]), 0u, 300 + t);      // the prior tiles in these columns have been drained
            tc_fence_after();
#pragma unroll 1
            for (int kc = 0; kc < kFwdChunks; ++kc)
                mma_chunk<4>(b, rp, tmem_base, d, d + 32, xh + 128 * kc, xl + 128 * kc, kc ? 1u : 0u, kc ? 1u : 0u, idesc, waited,
                             9 + kFwdChunks * t + kc, pg_trace_mma);
            mma_commit_acc(b, 3 + t);
        }
    } else {
        const uint32_t qh = desc_lo(B.dq_hi), ql = desc_lo(B.dq_lo);
#pragma unroll 1
        for (int p = 0; p < kBwdParts; ++p) {
            const uint32_t d = acc + 64 * p;
#pragma unroll 1
            for (int c = bwd_part_begin(p); c < bwd_part_begin(p + 1); ++c)
                mma_chunk<4>(b, rp, tmem_base, d, d + 32, qh + 128 * c, ql + 128 * c, c == bwd_part_begin(p) ? 0u : 1u,
                             c == bwd_part_begin(p) ? 0u : 1u, idesc, waited, c, pg_trace_mma);
            mma_commit_acc(b, p);
        }
    }
#if defined(SMPLB200_PHASE_CLOCKS)
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_phase_clocks[FWD ? 26 : 30], (unsigned long long)waited);
#endif
}

// ---- epilogue helpers ------------------------------------------------------------------------------------------------------
// this thread's 32 accumulator columns (16 samples of CTA 0, then 16 of CTA 1) of one accumulator
__device__ __forceinline__ void acc_load32(uint32_t taddr, float* v) {
    tmem_ld32(taddr, v);
    tmem_ld_wait();
}
__device__ __forceinline__ void acc_wait(Bars* b, int idx) {
    pg_wait(smem_u32(&b->acc_full[idx]), 0u, 400 + idx);
    tc_fence_after();
}

}  // namespace pg
}  // namespace smplb200
