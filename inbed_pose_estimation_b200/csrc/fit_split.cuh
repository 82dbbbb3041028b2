// SMPLify.__call__ (reference smplify/smplify.py:40-136) for SMALL batches (the reference's own operating point is
// --batch_size 32, README.md:33-35; trainer.py:709-715): a tile of 4 samples is fitted by a CLUSTER of C = 2, 4 or 8 CTAs.
//
// A 4-sample tile on one SM spends two thirds of an iteration streaming the 1.4 MB of GEMM constants (folded joint model
// forward and transpose, 8 prior precisions) from L2 at the SM's ingest rate, while most of the chip idles.  Here every CTA of
// the cluster holds the WHOLE tile state and runs the cheap per-sample phases redundantly (bit-identical in every CTA), but
// computes only its 1 / C of the OUTPUT ROWS of the three GEMMs - with the reduction dimension split over all of its threads -
// and stores the reduced rows into the shared memory of every CTA of the cluster (distributed shared memory,
// st.shared::cluster).  Two cluster barriers per iteration:
//
//   pose features, rest joints                                                            all CTAs, redundantly
//   prior quadratic forms + folded GEMM forward: own rows -> Pd / Q of every CTA           split     || chain forward sweep (last 1-2 warps)
//   -- cluster barrier (the chain warps arrive before their sweep; the prior selection runs beside its tail) --
//   prior selection, 49 output joints, projection + GMoF, joint backward                   all CTAs, redundantly
//   picked-vertex backward, folded GEMM backward: own rows of dL/dx -> landing rows of every CTA    || chain backward sweep
//   -- cluster barrier (again the chain warps arrive first) --
//   Rodrigues backward + Adam                                                              all CTAs, redundantly
//
// Sums over the reduction slices are added in a fixed order (deterministic; not the order of the 4-sample tile kernel: the
// two agree to rounding, tests/test_gpu_smplify.py).  CTA 0 of the cluster writes the results.
#pragma once
#include "fit_driver.cuh"

namespace smplb200 {

constexpr int kSplitS = 4;                         // samples per cluster

// Work split of the three GEMMs for a cluster of C CTAs: (output items per CTA) x (reduction slices) <= GT = 384 - 32 CW GEMM
// threads, the last CW warps run the kinematic-chain sweeps beside them (per half of the iteration: GTF / CWF, GTB / CWB); every slice length is a multiple of 4, at least 8
// (stream_gemm)
template <int C> struct SplitPlan;
template <> struct SplitPlan<8> {
    static constexpr int FQ = 22, FS = 14, FR = 16;      // forward: 22 column quads x 14 slices of 16 of the 224 x rows
    static constexpr int PI = 18, PS = 9, PR = 8;        // prior:   one component (18 column quads) x 9 slices of 8 of the 72 rows
    static constexpr int BQ = 7, BS = 44, BR = 16;       // backward: 7 column quads x 44 slices of 16 of the 704 dQ rows
    static constexpr int CWF = 2, GTF = kFitTileThreads - 32 * CWF;     // chain warps / GEMM threads of the forward half
    static constexpr int CWB = 2, GTB = kFitTileThreads - 32 * CWB;     // ... of the backward half
};
template <> struct SplitPlan<4> {
    static constexpr int FQ = 44, FS = 7, FR = 32;
    static constexpr int PI = 36, PS = 6, PR = 12;
    static constexpr int BQ = 14, BS = 22, BR = 32;
    static constexpr int CWF = 2, GTF = kFitTileThreads - 32 * CWF;     // chain warps / GEMM threads of the forward half
    static constexpr int CWB = 2, GTB = kFitTileThreads - 32 * CWB;     // ... of the backward half
};
template <> struct SplitPlan<2> {
    static constexpr int FQ = 88, FS = 4, FR = 56;
    static constexpr int PI = 72, PS = 3, PR = 24;
    static constexpr int BQ = 28, BS = 11, BR = 64;
    static constexpr int CWF = 1, GTF = kFitTileThreads - 32 * CWF;     // 88 column quads x 4 slices need 352 GEMM threads
    static constexpr int CWB = 2, GTB = kFitTileThreads - 32 * CWB;
};

// tile state of fit_tile.cuh (S = 4) + the regions the split needs
struct SplitLayout : TileLayout<kSplitS> {
    using Base = TileLayout<kSplitS>;
    static constexpr int PD = (Base::SMEM_FLOATS + 3) / 4 * 4;             // [8 x 72][4]  Psym (bp - mean) of all components
    static constexpr int DXL = PD + kGauss * kPriorPad * kSplitS;          // [224][4]     dL/dx as it lands from the cluster
    static constexpr int SCR_F = DXL + kXPad * kSplitS;                    // slice partials of the forward / backward GEMM
    static constexpr int SCR_P = SCR_F + 5632;                             // slice partials of the prior GEMM
    static constexpr int TREE = SCR_P + 5184;                              // ChainTree: the kinematic tree's index tables
    static constexpr int SMEM_FLOATS = TREE + 32;
    SB_HD static int pd(int g, int i, int s) { return PD + (g * kPriorPad + i) * kSplitS + s; }
};

#if defined(SMPLB200_PHASE_CLOCKS)
#define SPLIT_CLK_BEGIN() const long long split_t0 = clock64()
#define SPLIT_CLK_END(slot, tid) do { if (blockIdx.x == 0 && (int)threadIdx.x == (tid)) g_phase_clocks[slot] += (unsigned long long)(clock64() - split_t0); } while (0)
#else
#define SPLIT_CLK_BEGIN() ((void)0)
#define SPLIT_CLK_END(slot, tid) ((void)0)
#endif

namespace split {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// every thread of every CTA of the cluster; release / acquire at cluster scope: the rows stored into peers are visible after it
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the two halves on their own (whole warps): a warp that stores nothing into its peers arrives first and works on
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <int NT>
__device__ __forceinline__ void gemm_threads_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

__device__ __forceinline__ void st_cluster4(uint32_t addr, const float4& v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// the 4 samples of one row (16-byte aligned) -> sm[offset .. offset + 3] of every CTA of the cluster: one 16-byte store per CTA
// (scalar stores were measured to cost ~1.4 cycles each on the SM-to-SM network: 5 120 per CTA and forward exchange)
template <int C>
__device__ __forceinline__ void broadcast4(float* sm, int offset, const float4& v) {
    const uint32_t local = smem_u32(sm + offset);
#if defined(SPLIT_NO_REMOTE)                      // timing discriminator (wrong results): rows stay in the producing CTA
    *reinterpret_cast<float4*>(sm + offset) = v;
    (void)local;
#else
#pragma unroll
    for (int r = 0; r < C; ++r) st_cluster4(mapa(local, (uint32_t)r), v);
#endif
}

// partial sums of one thread's 4 columns x 4 samples -> scr[(slice * ncols + col) * 4 + s]
__device__ __forceinline__ void store_partial(const TileAcc<2>& acc, float* scr, int slice, int ncols, int col0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) acc.store(scr + ((size_t)slice * ncols + col0 + c) * kSplitS, c);
}
// fixed-order sum over the slices of one column (4 samples)
template <int NSLICE>
__device__ __forceinline__ float4 sum_slices(const float* scr, int ncols, int col) {
    const float4* p = reinterpret_cast<const float4*>(scr) + col;
    float4 a = p[0];
#pragma unroll 4
    for (int k = 1; k < NSLICE; ++k) {
        const float4 b = p[(size_t)k * ncols];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
}
}  // namespace split

// The kinematic-chain sweeps of fit_tile.cuh (ph_chain_forward / ph_chain_backward) with one work item per ROW of a joint's 3 x 4
// transform instead of per joint: a third of the instruction sequence per item, three times the items.  With 4 samples a tree
// level has at most 5 x 4 items - the sweeps are bound by the length of one item's dependent instruction sequence, not by
// throughput - so rows on two warps cut the sweep latency, which the split GEMMs no longer hide.  Every element is computed by
// the same expression as in fit_tile.cuh: bit-identical results.
// The tree's index tables in shared memory: every level of a sweep starts with a chain of dependent table look-ups
// (level -> joint -> parent / children), ~30 cycles each from shared memory against ~100 from the kernel-parameter bank.
struct ChainTree {
    int8_t parents[kJoints];
    uint8_t level_order[kJoints], level_start[kMaxLevels + 1], child_start[kJoints + 1], child_list[kJoints];
    int32_t num_levels;
};
static_assert(sizeof(ChainTree) <= 32 * sizeof(float), "tree tables region");
__device__ __forceinline__ void stage_chain_tree(const ModelView& M, ChainTree* t) {
    FOR_ITEMS(i, kJoints) { t->parents[i] = M.parents[i]; t->level_order[i] = M.level_order[i]; t->child_list[i] = M.child_list[i]; }
    FOR_ITEMS(i, kJoints + 1) t->child_start[i] = M.child_start[i];
    FOR_ITEMS(i, kMaxLevels + 1) t->level_start[i] = M.level_start[i];
    if (threadIdx.x == 0) t->num_levels = M.num_levels;
}

// one row r of joint j's world transform (parent p, -1 = root) for sample s
template <int S, class L>
__device__ __forceinline__ void chain_forward_row(float* sm, int j, int p, int r, int s) {
    float G[4];
    const float Jx = sm[L::JR + (3 * j + 0) * S + s], Jy = sm[L::JR + (3 * j + 1) * S + s], Jz = sm[L::JR + (3 * j + 2) * S + s];
    if (p < 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) G[c] = sm[L::RM + (j * 9 + r * 3 + c) * S + s];
        G[3] = (r == 0) ? Jx : ((r == 1) ? Jy : Jz);
    } else {
        float Rl[9], Gp[4];
#pragma unroll
        for (int e = 0; e < 9; ++e) Rl[e] = sm[L::RM + (j * 9 + e) * S + s];
#pragma unroll
        for (int e = 0; e < 4; ++e) Gp[e] = sm[L::GW + (p * 12 + r * 4 + e) * S + s];
        const float dx = Jx - sm[L::JR + (3 * p + 0) * S + s], dy = Jy - sm[L::JR + (3 * p + 1) * S + s],
                    dz = Jz - sm[L::JR + (3 * p + 2) * S + s];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[c] = Gp[0] * Rl[c] + Gp[1] * Rl[3 + c] + Gp[2] * Rl[6 + c];
        G[3] = Gp[0] * dx + Gp[1] * dy + Gp[2] * dz + Gp[3];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) sm[L::GW + (j * 12 + r * 4 + e) * S + s] = G[e];
    sm[L::AT + (3 * j + r) * S + s] = G[3] - (G[0] * Jx + G[1] * Jy + G[2] * Jz);
}

// Level by level.  A thread's first item of the NEXT level (joint, parent: three dependent table look-ups that do not depend on
// this level's results) is looked up before this level's arithmetic, so that it is off the level-to-level critical path.
template <int S, class L>
__device__ __forceinline__ void ph_chain_forward_rows(const ChainTree& M, float* sm, const Grp g) {
    const int levels = M.num_levels, s = g.tid % S, r = (g.tid / S) % 3, slot = g.tid / (3 * S);
    int first = M.level_start[0], cnt = M.level_start[1] - first;
    int j0 = -1, p0 = -1;
    if (slot < cnt) { j0 = M.level_order[first + slot]; p0 = M.parents[j0]; }
    for (int lev = 0; lev < levels; ++lev) {
        int nfirst = 0, ncnt = 0, nj = -1, np = -1;
        if (lev + 1 < levels) {
            nfirst = M.level_start[lev + 1];
            ncnt = M.level_start[lev + 2] - nfirst;
            if (slot < ncnt) { nj = M.level_order[nfirst + slot]; np = M.parents[nj]; }
        }
        if (j0 >= 0) chain_forward_row<S, L>(sm, j0, p0, r, s);
        for (int it = g.tid + g.nt; it < cnt * 3 * S; it += g.nt) {              // further passes of a wide level / a narrow group
            const int j = M.level_order[first + it / (3 * S)];
            chain_forward_row<S, L>(sm, j, M.parents[j], (it / S) % 3, it % S);
        }
        grp_sync(g);
        first = nfirst; cnt = ncnt; j0 = nj; p0 = np;
    }
}

// row r of parent p's accumulated dL/dG and entry r of its rest-joint gradient, gathered from its children [c0, c1)
template <int S, class L>
__device__ __forceinline__ void chain_backward_row(const ChainTree& M, float* sm, int p, int c0, int c1, int r, int s) {
    float dGp[4], Gp[3];
#pragma unroll
    for (int e = 0; e < 4; ++e) dGp[e] = sm[L::DG + (p * 12 + r * 4 + e) * S + s];
#pragma unroll
    for (int k = 0; k < 3; ++k) Gp[k] = sm[L::GW + (p * 12 + k * 4 + r) * S + s];            // column r of G_p^R
    float dJp = sm[L::DJ + (3 * p + r) * S + s];
    const float Jpx = sm[L::JR + (3 * p + 0) * S + s], Jpy = sm[L::JR + (3 * p + 1) * S + s], Jpz = sm[L::JR + (3 * p + 2) * S + s];
    for (int ci = c0; ci < c1; ++ci) {
        const int i = M.child_list[ci];
        float dGi[4], Ri[9];
#pragma unroll
        for (int e = 0; e < 4; ++e) dGi[e] = sm[L::DG + (i * 12 + r * 4 + e) * S + s];
#pragma unroll
        for (int e = 0; e < 9; ++e) Ri[e] = sm[L::RM + (i * 9 + e) * S + s];
        const float t0 = sm[L::DG + (i * 12 + 3) * S + s], t1 = sm[L::DG + (i * 12 + 7) * S + s], t2 = sm[L::DG + (i * 12 + 11) * S + s];
        const float rel[3] = {sm[L::JR + (3 * i + 0) * S + s] - Jpx, sm[L::JR + (3 * i + 1) * S + s] - Jpy,
                              sm[L::JR + (3 * i + 2) * S + s] - Jpz};
#pragma unroll
        for (int c = 0; c < 3; ++c)
            dGp[c] += dGi[0] * Ri[c * 3 + 0] + dGi[1] * Ri[c * 3 + 1] + dGi[2] * Ri[c * 3 + 2] + dGi[3] * rel[c];
        dGp[3] += dGi[3];
        dJp -= Gp[0] * t0 + Gp[1] * t1 + Gp[2] * t2;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) sm[L::DG + (p * 12 + r * 4 + e) * S + s] = dGp[e];
    sm[L::DJ + (3 * p + r) * S + s] = dJp;
}

template <int S, class L>
__device__ __forceinline__ void ph_chain_backward_rows(const ChainTree& M, float* sm, const Grp g) {
    const int levels = M.num_levels, s = g.tid % S, r = (g.tid / S) % 3, slot = g.tid / (3 * S);
    if (levels < 2) return;
    int first = M.level_start[levels - 2], cnt = M.level_start[levels - 1] - first;
    int p0 = -1, c0 = 0, c1 = 0;
    if (slot < cnt) { p0 = M.level_order[first + slot]; c0 = M.child_start[p0]; c1 = M.child_start[p0 + 1]; }
    for (int lev = levels - 2; lev >= 0; --lev) {
        int nfirst = 0, ncnt = 0, np = -1, nc0 = 0, nc1 = 0;
        if (lev > 0) {
            nfirst = M.level_start[lev - 1];
            ncnt = M.level_start[lev] - nfirst;
            if (slot < ncnt) { np = M.level_order[nfirst + slot]; nc0 = M.child_start[np]; nc1 = M.child_start[np + 1]; }
        }
        if (p0 >= 0 && c0 != c1) chain_backward_row<S, L>(M, sm, p0, c0, c1, r, s);
        for (int it = g.tid + g.nt; it < cnt * 3 * S; it += g.nt) {
            const int p = M.level_order[first + it / (3 * S)], a = M.child_start[p], b = M.child_start[p + 1];
            if (a != b) chain_backward_row<S, L>(M, sm, p, a, b, (it / S) % 3, it % S);
        }
        grp_sync(g);
        first = nfirst; cnt = ncnt; p0 = np; c0 = nc0; c1 = nc1;
    }
}
// ... and its per-joint tail (dL/dR_j over RM, rest-joint gradient), independent of the tree order: the whole tile runs it
template <int S, class L>
__device__ __forceinline__ void ph_chain_backward_finish_rows(const ChainTree& M, float* sm, const Grp g) {
    FOR_ITEMS_G(it, kJoints * 3 * S, g) {
        const int s = it % S, r = (it / S) % 3, j = it / (3 * S), p = M.parents[j];
        if (p < 0) {
#pragma unroll
            for (int c = 0; c < 3; ++c) sm[L::RM + (j * 9 + r * 3 + c) * S + s] = sm[L::DG + (j * 12 + r * 4 + c) * S + s];
            sm[L::DJ + (3 * j + r) * S + s] += sm[L::DG + (j * 12 + r * 4 + 3) * S + s];
        } else {
            float dGi[12], Gp[3];
#pragma unroll
            for (int e = 0; e < 12; ++e) dGi[e] = sm[L::DG + (j * 12 + e) * S + s];
#pragma unroll
            for (int k = 0; k < 3; ++k) Gp[k] = sm[L::GW + (p * 12 + k * 4 + r) * S + s];            // column r of G_p^R
#pragma unroll
            for (int c = 0; c < 3; ++c)   // row r of dR_j = G_p^R^T dG_j^R
                sm[L::RM + (j * 9 + r * 3 + c) * S + s] = Gp[0] * dGi[0 + c] + Gp[1] * dGi[4 + c] + Gp[2] * dGi[8 + c];
            sm[L::DJ + (3 * j + r) * S + s] += Gp[0] * dGi[3] + Gp[1] * dGi[7] + Gp[2] * dGi[11];
        }
    }
}

// This CTA's rows of Pd = Psym (bp - mean) (all 8 components) and of Q = Cf^T x, into every CTA of the cluster.  The GEMM
// threads (SplitPlan::GT) call it; the caller follows with a cluster barrier.
template <int C, class L>
__device__ __forceinline__ void split_forward_gemms(const ModelView& M, const SmallConsts& Cn, float* sm, uint32_t rank) {
    using PL = SplitPlan<C>;
    constexpr int S = kSplitS, NQ4 = kQPad / 4, IQ = kPriorPad / 4;
    static_assert(PL::FQ * C == NQ4 && PL::FS * PL::FR == kXPad && PL::FQ * PL::FS <= PL::GTF, "forward split");
    static_assert(PL::PI * C == kGauss * IQ && PL::PS * PL::PR == kPriorPad && PL::PI * PL::PS <= PL::GTF, "prior split");
    static_assert(PL::FS * PL::FQ * 16 <= 5632 && PL::PS * PL::PI * 16 <= 5184, "scratch sizes");
    const int t = (int)threadIdx.x;
    float* scr_p = sm + L::SCR_P;
    float* scr_f = sm + L::SCR_F;
    if (t < PL::PI * PL::PS) {
        const int il = t % PL::PI, ks = t / PL::PI, gi = (int)rank * PL::PI + il, g = gi / IQ, iq = gi % IQ;
        TileAcc<2> acc;
        acc.clear();
        stream_gemm<PL::PR, IQ, S>(acc, reinterpret_cast<const float4*>(M.gmm_prec + (size_t)g * kPriorPad * kPriorPad) + iq + (size_t)ks * PL::PR * IQ,
                                   sm + L::POSE + (3 + ks * PL::PR) * S);
        split::store_partial(acc, scr_p, ks, PL::PI * 4, 4 * il);
    }
    if (t < PL::FQ * PL::FS) {
        const int ql = t % PL::FQ, ks = t / PL::FQ, cq = (int)rank * PL::FQ + ql;
        TileAcc<2> acc;
        acc.clear();
        stream_gemm<PL::FR, NQ4, S>(acc, reinterpret_cast<const float4*>(M.Cf) + cq + (size_t)ks * PL::FR * NQ4, sm + L::XT + ks * PL::FR * S);
        split::store_partial(acc, scr_f, ks, PL::FQ * 4, 4 * ql);
    }
    split::gemm_threads_sync<PL::GTF>();
    for (int col = t; col < PL::PI * 4; col += PL::GTF) {
        const int gi4 = (int)rank * PL::PI * 4 + col;                                 // = g * 72 + i (18 quads of 4 per component)
        float4 v = split::sum_slices<PL::PS>(scr_p, PL::PI * 4, col);
        const float pm = Cn.pmean[gi4];
        v.x -= pm; v.y -= pm; v.z -= pm; v.w -= pm;
        split::broadcast4<C>(sm, L::PD + gi4 * S, v);
    }
    for (int col = t; col < PL::FQ * 4; col += PL::GTF)
        split::broadcast4<C>(sm, L::q((int)rank * PL::FQ * 4 + col, 0), split::sum_slices<PL::FS>(scr_f, PL::FQ * 4, col));
}

// This CTA's rows of dL/dx = Cf dQ into the landing rows (DXL) of every CTA of the cluster
template <int C, class L>
__device__ __forceinline__ void split_backward_gemm(const ModelView& M, float* sm, uint32_t rank) {
    using PL = SplitPlan<C>;
    constexpr int S = kSplitS, MQ = kXPad / 4;
    static_assert(PL::BQ * C == MQ && PL::BS * PL::BR == kQPad && PL::BQ * PL::BS <= PL::GTB, "backward split");
    static_assert(PL::BS * PL::BQ * 16 <= 5632, "scratch size");
    const int t = (int)threadIdx.x;
    float* scr = sm + L::SCR_F;
    if (t < PL::BQ * PL::BS) {
        const int ml = t % PL::BQ, ns = t / PL::BQ, mq = (int)rank * PL::BQ + ml;
        TileAcc<2> acc;
        acc.clear();
        stream_gemm<PL::BR, MQ, L::LDQ>(acc, reinterpret_cast<const float4*>(M.CfT) + (size_t)ns * PL::BR * MQ + mq, sm + L::QT + ns * PL::BR * L::LDQ);
        split::store_partial(acc, scr, ns, PL::BQ * 4, 4 * ml);
    }
    split::gemm_threads_sync<PL::GTB>();
    for (int col = t; col < PL::BQ * 4; col += PL::GTB)
        split::broadcast4<C>(sm, L::DXL + ((int)rank * PL::BQ * 4 + col) * S, split::sum_slices<PL::BS>(scr, PL::BQ * 4, col));
}

// Forward through the folded model for the tile's current parameters, up to (not including) the 49 output joints.  The chain
// warps arrive at the cluster barrier BEFORE their sweep (they store nothing into the peers); the GEMM threads, once the
// cluster's rows have landed, run the prior selection while the sweep - the longer of the two jobs - is still finishing.
template <int C, class L>
__device__ __forceinline__ void split_forward(const ModelView& M, const SmallConsts& Cn, float* sm, uint32_t rank, bool root_identity,
                                              bool with_prior) {
    constexpr int S = kSplitS;
    using PL = SplitPlan<C>;
    PHASE_BEGIN();
    ph_pose_features<S, L>(sm, true, root_identity);
    ph_rest_joints<S, L>(Cn, sm);
    TILE_SYNC();
    PHASE_MARK(0);
    if ((int)threadIdx.x >= PL::GTF) {
        SPLIT_CLK_BEGIN();
        split::cluster_arrive();
        ph_chain_forward_rows<S, L>(*reinterpret_cast<const ChainTree*>(sm + L::TREE), sm, Grp{(int)threadIdx.x - PL::GTF, 32 * PL::CWF, PL::CWF == 1 ? 1 : 2});
        SPLIT_CLK_END(10, PL::GTF);
        split::cluster_wait();
    } else {
        split_forward_gemms<C, L>(M, Cn, sm, rank);
        PHASE_MARK(1);
        split::cluster_sync();
        PHASE_MARK(12);
        if (with_prior) ph_prior_select<S, L, 8>(M, Cn, sm, kPosePriorW2, kAnglePriorW2, kShapePriorW2, Grp{(int)threadIdx.x, PL::GTF, 3});
        PHASE_MARK(13);
    }
    TILE_SYNC();
}

// the whole fit of one 4-sample tile by a cluster of C CTAs of kFitThreads threads; `first` = first sample of the tile
template <int C>
__device__ void fit_split_tile(const ModelView& M, const FitParams& Pin, int first, float* sm) {
    using L = SplitLayout;
    constexpr int S = kSplitS;
    const uint32_t rank = split::cluster_rank();
    FitParams P = Pin;
    if (rank != 0) P.loss_trace = nullptr;                    // one writer per sample
    AdamScalars* adam_tab = reinterpret_cast<AdamScalars*>(sm + L::ADAMTAB);
    const SmallConsts Cn = stage_small_consts<S, L>(M, sm);
    stage_chain_tree(M, reinterpret_cast<ChainTree*>(sm + L::TREE));
    FOR_ITEMS(t, (P.num_iters < kMaxIters ? P.num_iters : kMaxIters)) adam_tab[t] = adam_scalars(P, t);
    FOR_ITEMS(it, S * 72) {
        const int s = it / 72, k = it % 72, b = first + s;
        sm[L::POSE + k * S + s] = (b < P.batch) ? P.init_pose[(size_t)b * 72 + k] : 0.f;
    }
    FOR_ITEMS(it, S * kBetas) {
        const int s = it / kBetas, k = it % kBetas, b = first + s;
        sm[L::BETA + k * S + s] = (b < P.batch) ? P.init_betas[(size_t)b * kBetas + k] : 0.f;
    }
    FOR_ITEMS(it, S * 3) {
        const int s = it / 3, k = it % 3, b = first + s;
        sm[L::CAM + k * S + s] = (b < P.batch) ? P.init_cam[(size_t)b * 3 + k] : (k == 2 ? 1.f : 0.f);
    }
    FOR_ITEMS(it, S * 2) {
        const int s = it / 2, k = it % 2, b = first + s;
        sm[L::CEN + k * S + s] = (b < P.batch) ? P.center[(size_t)b * 2 + k] : 0.f;
    }
    FOR_ITEMS(it, S * 147) {
        const int s = it / 147, k = it % 147, b = first + s;
        sm[L::KP + k * S + s] = (b < P.batch) ? P.keypoints[(size_t)b * 147 + k] : 0.f;
    }
    // every CTA has read the keypoints before any CTA zeroes confidences in place; also: all CTAs of the cluster are running
    // before the first store into a peer's shared memory
    split::cluster_sync();
    if (P.zero_conf_first) { tile_zero_ignored_conf<S, L>(M, P, first, sm); TILE_SYNC(); }

    if (P.num_iters > 0) {
        // ---- stage 1: global orientation + camera translation (every CTA, redundantly) ---------
        split_forward<C, L>(M, Cn, sm, rank, /*root_identity=*/true, false);
        ph_output_joints<S, L>(M, Cn, sm);
        TILE_SYNC();
        stage1_camera<S, L>(M, P, first, sm);
        TILE_SYNC();
        tile_zero_ignored_conf<S, L>(M, P, first, sm);
        zero_rows<S, L>(sm, L::ADM, 2 * kParams);
        TILE_SYNC();

        // ---- stage 2 ---------------------------------------------------------------------------
        PHASE_BEGIN();
        for (int it = 0; it < P.num_iters; ++it) {
            split_forward<C, L>(M, Cn, sm, rank, false, true);
            PHASE_MARK(2);
            ph_output_joints<S, L>(M, Cn, sm);
            TILE_SYNC();
            PHASE_MARK(3);
            ph_reprojection<S, L>(sm, P.focal, kSigma2, true);
            zero_rows<S, L>(sm, L::DG, 288);
            TILE_SYNC();
            if (P.loss_trace) {
                FOR_ITEMS(s, S) {
                    const int b = first + s;
                    float a = 0.f;
                    for (int o = 0; o < kOut; ++o) a += sm[L::LOSSJ + o * S + s];
                    a = ((a + sm[L::LOSSJ + 49 * S + s]) + sm[L::LOSSJ + 50 * S + s]) + sm[L::LOSSJ + 51 * S + s];
                    if (b < P.batch) P.loss_trace[(size_t)(P.num_iters + it) * P.batch + b] = a;
                }
            }
            PHASE_MARK(4);
            ph_joint_backward<S, L>(M, Cn, sm);
            TILE_SYNC();
            PHASE_MARK(5);
            // the picked-vertex backward only adds dQ rows (it reads the source gradients and the transforms): the chain warps
            // skip it and start their reverse sweep - the longest job of this half - on what the joint backward left
            using PL = SplitPlan<C>;
            if ((int)threadIdx.x >= PL::GTB) {
                SPLIT_CLK_BEGIN();
                split::cluster_arrive();                           // nothing of the sweep goes to a peer
                ph_chain_backward_rows<S, L>(*reinterpret_cast<const ChainTree*>(sm + L::TREE), sm, Grp{(int)threadIdx.x - PL::GTB, 32 * PL::CWB, PL::CWB == 1 ? 1 : 2});
                SPLIT_CLK_END(11, PL::GTB);
                split::cluster_wait();
            } else {
                static_assert(kPicks * S <= PL::GTB, "one picked-vertex item per GEMM thread");
                ph_pick_backward<S, L>(M, Cn, sm);                 // items < GT: the chain warps own none
                split::gemm_threads_sync<PL::GTB>();
                PHASE_MARK(6);
                split_backward_gemm<C, L>(M, sm, rank);
                PHASE_MARK(7);
                split::cluster_sync();
                PHASE_MARK(8);
                for (int i = (int)threadIdx.x; i < kXPad * S; i += PL::GTB) sm[L::XT + i] = sm[L::DXL + i];      // beside the sweep's tail
            }
            TILE_SYNC();
            ph_chain_backward_finish_rows<S, L>(*reinterpret_cast<const ChainTree*>(sm + L::TREE), sm, grp_tile());
            TILE_SYNC();
            const AdamScalars sc = (it < kMaxIters) ? adam_tab[it] : adam_scalars(P, it);
            tile_adam_step<S, L>(Cn, P, sm, sc);
            TILE_SYNC();
            PHASE_MARK(9);
        }
    }

    // ---- final forward; CTA 0 writes the results ------------------------------------------------
    split_forward<C, L>(M, Cn, sm, rank, false, false);     // holds the last cluster barrier: no store into a peer after it
    ph_output_joints<S, L>(M, Cn, sm);
    TILE_SYNC();
    if (rank == 0) tile_write_outputs<S, L>(P, first, sm);
}
}  // namespace smplb200
