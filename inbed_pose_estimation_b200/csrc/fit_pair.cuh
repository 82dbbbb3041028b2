// SMPLify.__call__ (reference smplify/smplify.py:40-136) for a PAIR of CTAs: two tiles of S samples, one per CTA of a 2-CTA
// cluster, run the per-sample phases of fit_tile.cuh on their own samples and share the three per-iteration GEMMs on the
// tensor cores (pair_gemm.cuh): the folded joint model forward (+ the 8-component prior quadratic forms in the same call)
// and its transpose.  Per iteration and CTA:
//
//   pose features, rest joints, B operands (x, body pose: raw fp32 = tf32 hi part, remainder = lo part)      all threads
//   -- cluster barrier --
//   forward call     generators | MMA issue | chain forward sweep (warp 13) | epilogue warps: prior tiles -> both CTAs,
//                    prior selection (arg-min, gradient) on the 128 epilogue threads while the forward tiles are still
//                    being multiplied, forward tiles -> Q of both CTAs
//   -- cluster barrier --
//   49 output joints, projection + GMoF, joint / picked-vertex backward (dQ hi + lo written in operand layout)  all threads
//   -- cluster barrier --
//   backward call    generators | MMA issue | chain backward sweep | epilogue: 4 partial accumulator pairs -> dx of both CTAs
//   -- cluster barrier --
//   Rodrigues backward + Adam                                                                                    all threads
#pragma once
#include "fit_driver.cuh"
#include "pair_gemm.cuh"

namespace smplb200 {

SB_HD float tf32_trunc_hd(float v) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#else
    union { float f; unsigned u; } c;
    c.f = v;
    c.u &= 0xFFFFE000u;
    return c.f;
#endif
}

// Shared-memory layout of one CTA of the pair.  The three tensor-core operand regions come first (the kernel aligns the
// base to 1024 bytes: SWIZZLE_128B atoms); everything else is the [index][S] state of fit_tile.cuh.
template <int S_>
struct PairLayout {
    static constexpr int S = S_;
    static_assert(S % 4 == 0 && S <= 16, "a CTA of the pair holds at most 16 samples (the MMA's N is 2 x 16)");
    static constexpr int kAtom = 16 * 32;
    static constexpr int XT = 0;                          // x (B operand hi, 7 atoms) | dx [224][S] | source-gradient scratch
    static constexpr int QT = XT + 7 * kAtom;             // Q, then dQ (B operand hi, 22 atoms)
    static constexpr int LO = QT + 22 * kAtom;            // dQ lo (22 atoms); during the forward call x lo | body pose hi | lo
    static constexpr int XLO = LO, BPH = LO + 7 * kAtom, BPL = BPH + 3 * kAtom;
    static constexpr int POSE = LO + 22 * kAtom;          // [72][S]
    static constexpr int BETA = POSE + 72 * S;
    static constexpr int CAM = BETA + 10 * S;
    static constexpr int CEN = CAM + 3 * S;
    static constexpr int KP = CEN + 2 * S;
    static constexpr int RM = KP + 147 * S;
    static constexpr int JR = RM + 216 * S;
    static constexpr int GW = JR + 72 * S;
    static constexpr int AT = GW + 288 * S;
    static constexpr int OUTJ = AT + 72 * S;              // OUTJ, DG, DJ and the 49 reprojection rows of LOSSJ are dead while the
    static constexpr int DG = OUTJ + 147 * S;             // forward call runs: together they hold Pd [8][69][S] (PD) for the
    static constexpr int DJ = DG + 288 * S;               // prior selection, so that Q can land in QT independently of it
    static constexpr int LOSSJ = DJ + 72 * S;
    static constexpr int GPR = LOSSJ + 52 * S;
    static constexpr int ADM = GPR + 69 * S;
    static constexpr int PD = OUTJ;
    static_assert(kGauss * kPriorDim <= 147 + 288 + 72 + 49, "Pd must end before the prior rows of LOSSJ");
    static constexpr int ADV = ADM + 82 * S;
    static constexpr int MISC = ADV + 82 * S;
    static constexpr int TOTAL = MISC + 16 * S;
    static constexpr int ADAMTAB = TOTAL;
    static constexpr int CONSTS = ADAMTAB + 2 * kMaxIters;
    static constexpr int BARS = (CONSTS + kSmallConstFloats + 1) / 2 * 2;          // 8-byte aligned
    static constexpr int SMEM_FLOATS = BARS + 40;
    static_assert(kXPad * S <= 7 * kAtom, "dx must fit in the XT region");
    static constexpr bool kPair = true;
    SB_HD static int pd(int g, int i, int s) { return PD + (g * kPriorDim + i) * S + s; }
    SB_HD static int b_off(int s, int k) { return (k >> 5) * kAtom + s * 32 + ((((k & 31) >> 2) ^ (s & 7)) << 2) + (k & 3); }
    SB_HD static int q(int n, int s) { return QT + b_off(s, n); }
    SB_HD static int x(int m, int s) { return XT + b_off(s, m); }
    // the tensor core reads the top 19 bits of the raw value as the hi part; the remainder goes to the lo operand
    SB_HD static void store_dq(float* sm, int n, int s, float v) {
        const int o = b_off(s, n);
        sm[QT + o] = v;
        sm[LO + o] = v - tf32_trunc_hd(v);
    }
    SB_HD static void store_x(float* sm, int m, int s, float v) {
        const int o = b_off(s, m);
        sm[XT + o] = v;
        sm[XLO + o] = v - tf32_trunc_hd(v);
    }
};

#if defined(__CUDACC__)

// Per-phase cycle counters of the pair kernel (profiling builds only, -DSMPLB200_PHASE_CLOCKS; tools/phase_clocks.py --pair).
// Slots 0-15: thread 0 of CTA 0 (a generator warp); 16-31: lane 0 of the MMA / epilogue / chain warps of CTA 0.
#if defined(SMPLB200_PHASE_CLOCKS)
#define PAIR_CLK_DECL() long long pclk_t0 = clock64(); const bool pclk_on = (blockIdx.x == 0 && (threadIdx.x & 31) == 0)
#define PAIR_CLK_RESET() do { pclk_t0 = clock64(); } while (0)
#define PAIR_CLK(i) do { if (pclk_on) { const long long t = clock64(); atomicAdd(&g_phase_clocks[i], (unsigned long long)(t - pclk_t0)); pclk_t0 = t; } } while (0)
#else
#define PAIR_CLK_DECL() ((void)0)
#define PAIR_CLK_RESET() ((void)0)
#define PAIR_CLK(i) ((void)0)
#endif

// body pose (POSE rows 3..71) as the B operand of the prior GEMM, K = 72 padded to 96 with zeros
template <int S, class L>
__device__ __forceinline__ void ph_prior_operand(float* sm) {
    FOR_ITEMS(it, 96 * S) {
        const int s = it % S, k = it / S;
        const float v = (k < kPriorDim) ? sm[L::POSE + (3 + k) * S + s] : 0.f;
        const int o = L::b_off(s, k);
        sm[L::BPH + o] = v;
        sm[L::BPL + o] = v - tf32_trunc_hd(v);
    }
}

// The 32 accumulator columns of a thread are the 16 samples of CTA 0 followed by the 16 of CTA 1: own / peer halves by
// predicated selects (a rank-dependent index into a register array would send it to local memory).
__device__ __forceinline__ void pair_split_halves(const float (&v)[32], uint32_t rank, float (&own)[16], float (&peer)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        own[i] = rank ? v[16 + i] : v[i];
        peer[i] = rank ? v[i] : v[16 + i];
    }
}

template <int S, class L>
__device__ __forceinline__ void pair_write_rows(float* sm, uint32_t peer_base, int off, const float (&own)[16], const float (&peer)[16]) {
    // S consecutive floats at float offset `off` of the tile state: own samples locally, the peer's samples into the peer CTA
#pragma unroll
    for (int s4 = 0; s4 < S / 4; ++s4) {
        *reinterpret_cast<float4*>(sm + off + 4 * s4) = make_float4(own[4 * s4], own[4 * s4 + 1], own[4 * s4 + 2], own[4 * s4 + 3]);
        pg::st_remote_v4(peer_base + (uint32_t)(off + 4 * s4) * 4u, make_float4(peer[4 * s4], peer[4 * s4 + 1], peer[4 * s4 + 2], peer[4 * s4 + 3]));
    }
}

// epilogue warps of the forward call
template <int S, class L>
__device__ __forceinline__ void pair_epilogue_forward(const ModelView& M, const SmallConsts& C, float* sm, pg::Bars* bars,
                                                      uint32_t rank, uint32_t tmem_base, bool do_select) {
    const int warp = threadIdx.x >> 5, qr = warp - pg::kEpiWarp0, r = 32 * qr + (threadIdx.x & 31);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * qr) << 16);
    const uint32_t peer_base = pg::mapa(tc::smem_u32(sm), rank ^ 1u);
    const uint32_t free_leader = pg::mapa(tc::smem_u32(&bars->acc_free[0]), 0);
    PG_TRACE_CALL_BEGIN();
    // prior tiles: Pd[g][i][s] = Psym (bp) - Psym mean, rows (g, i) of the PD region of the sample's CTA
    for (int t = 0; t < pg::kPriorTiles; ++t) {
        pg::acc_wait(bars, t);
        if (qr == 0) PG_TRACE_EVT(6, t);
        float v[32];
        pg::acc_load32(lane_addr + pg::kAccCol + 32 * t, v);
        if (t >= 1) {                                                   // tiles 0+1, then tile 2, have left TMEM: forward tiles may overwrite
            tc::tc_fence_before();
            __syncwarp();
            if ((threadIdx.x & 31) == 0) pg::mbar_arrive_remote(free_leader + (t - 1) * 8);
        }
        const int R = 256 * t + 128 * (int)rank + r, g = R / kPriorPad, i = R % kPriorPad;
        if (g < kGauss && i < kPriorDim) {
            const float pm = C.pmean[R];
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] -= pm;
            float own[16], peer[16];
            pair_split_halves(v, rank, own, peer);
            pair_write_rows<S, L>(sm, peer_base, L::pd(g, i, 0), own, peer);
        }
    }
    if (qr == 0) PG_TRACE_EVT(6, 8);
    pg::epi_pair_sync(bars, rank);                                  // both CTAs' Pd rows are complete
    if (qr == 0) PG_TRACE_EVT(6, 9);
    if (do_select)
        ph_prior_select<S, L>(M, C, sm, kPosePriorW2, kAnglePriorW2, kShapePriorW2, Grp{(int)threadIdx.x - 32 * pg::kEpiWarp0, 128, 2});
    if (qr == 0) PG_TRACE_EVT(6, 10);
    tc::tc_fence_after();
    for (int t = 0; t < pg::kFwdTiles; ++t) {
        pg::acc_wait(bars, 3 + t);
        if (qr == 0) PG_TRACE_EVT(6, 3 + t);
        float a[32], c[32];
        tc::tmem_ld32(lane_addr + pg::kAccCol + 64 * t, a);
        tc::tmem_ld32(lane_addr + pg::kAccCol + 64 * t + 32, c);
        tc::tmem_ld_wait();
        const int n = 256 * t + 128 * (int)rank + r;
        if (n < kQPad) {
#pragma unroll
            for (int k = 0; k < 32; ++k) a[k] += c[k];
            float own[16], peer[16];
            pair_split_halves(a, rank, own, peer);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int o = L::q(n, s);
                sm[o] = own[s];
                pg::st_remote_f32(peer_base + (uint32_t)o * 4u, peer[s]);
            }
        }
    }
    if (qr == 0) PG_TRACE_EVT(6, 12);
    tc::tc_fence_before();
}

// epilogue warps of the backward call: dx[m][s] = sum of the four partial accumulator pairs, plain [224][S] rows in XT
template <int S, class L>
__device__ __forceinline__ void pair_epilogue_backward(float* sm, pg::Bars* bars, uint32_t rank, uint32_t tmem_base) {
    const int warp = threadIdx.x >> 5, qr = warp - pg::kEpiWarp0, r = 32 * qr + (threadIdx.x & 31);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * qr) << 16);
    const uint32_t peer_base = pg::mapa(tc::smem_u32(sm), rank ^ 1u);
    float sum[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) sum[i] = 0.f;
    for (int p = 0; p < pg::kBwdParts; ++p) {
        pg::acc_wait(bars, p);
        float a[32], c[32];
        tc::tmem_ld32(lane_addr + pg::kAccCol + 64 * p, a);
        tc::tmem_ld32(lane_addr + pg::kAccCol + 64 * p + 32, c);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sum[i] += a[i] + c[i];
    }
    const int m = 128 * (int)rank + r;
    if (m < kXPad) {
        float own[16], peer[16];
        pair_split_halves(sum, rank, own, peer);
        pair_write_rows<S, L>(sm, peer_base, L::XT + m * S, own, peer);
    }
    tc::tc_fence_before();
}

// whole-cluster barrier between the per-sample phases and a GEMM call: orders generic-proxy shared-memory writes (local and
// remote) before the async-proxy reads of the MMAs and before the peer's accesses.  `bars` (before a call): the engine's
// barriers are re-initialised first - every wait of the previous call has returned by now.
__device__ __forceinline__ void pair_phase_barrier(pg::Bars* bars = nullptr) {
    if (bars && threadIdx.x == 0) {
        pg::bars_init(bars, true);
        if (blockIdx.x == 0) ++pg::g_pg_call_index;
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    pg::cluster_arrive();
    pg::cluster_wait();
    tc::tc_fence_after();
}

template <int S, class L>
__device__ __forceinline__ void pair_forward(const ModelView& M, const SmallConsts& C, const pg::Consts& K, float* sm, pg::Bars* bars,
                                             const pg::BOperands& B, uint32_t rank, uint32_t tmem_base,
                                             bool root_identity, bool do_select) {
    PAIR_CLK_DECL();
    const int warp = threadIdx.x >> 5;
    ph_pose_features<S, L>(sm, true, root_identity);
    ph_rest_joints<S, L>(C, sm);
    ph_prior_operand<S, L>(sm);
    if (warp == 0) PAIR_CLK(0);
    pair_phase_barrier(bars);
    if (warp == 0) PAIR_CLK(1);
    PAIR_CLK_RESET();
    if (warp < pg::kEpiWarp0) { pg::generator_run<true>(K, bars, rank, tmem_base); if (warp == 0) PAIR_CLK(2); if (warp == 4) PAIR_CLK(16); }
    else if (warp < pg::kMmaWarp) { pair_epilogue_forward<S, L>(M, C, sm, bars, rank, tmem_base, do_select); if (warp == pg::kEpiWarp0) PAIR_CLK(17); }
    else if (warp == pg::kMmaWarp) { if (rank == 0) pg::mma_run<true>(bars, B, tmem_base); PAIR_CLK(18); }
    else { ph_chain_forward<S, L>(M, sm, Grp{(int)threadIdx.x - 32 * pg::kChainWarp, 32, 1}); PAIR_CLK(19); }
    pair_phase_barrier();
    if (warp == 0) PAIR_CLK(3);
}

template <int S, class L>
__device__ __forceinline__ void pair_backward(const ModelView& M, const pg::Consts& K, float* sm, pg::Bars* bars,
                                              const pg::BOperands& B, uint32_t rank, uint32_t tmem_base) {
    PAIR_CLK_DECL();
    const int warp = threadIdx.x >> 5;
    pair_phase_barrier(bars);
    if (warp == 0) PAIR_CLK(8);
    PAIR_CLK_RESET();
    if (warp < pg::kEpiWarp0) { pg::generator_run<false>(K, bars, rank, tmem_base); if (warp == 0) PAIR_CLK(9); if (warp == 4) PAIR_CLK(20); }
    else if (warp < pg::kMmaWarp) { pair_epilogue_backward<S, L>(sm, bars, rank, tmem_base); if (warp == pg::kEpiWarp0) PAIR_CLK(21); }
    else if (warp == pg::kMmaWarp) { if (rank == 0) pg::mma_run<false>(bars, B, tmem_base); PAIR_CLK(22); }
    else { ph_chain_backward<S, L>(M, sm, Grp{(int)threadIdx.x - 32 * pg::kChainWarp, 32, 1}); PAIR_CLK(23); }
    pair_phase_barrier();
    if (warp == 0) PAIR_CLK(10);
}

// The whole two-stage fit of this CTA's S samples (rows first .. first + S of the batch); `sm` is 1024-byte aligned.
template <int S>
__device__ void fit_pair_tile(const ModelView& M, const FitParams& P, int first, float* sm, uint32_t rank) {
    using L = PairLayout<S>;
    pg::Bars* bars = reinterpret_cast<pg::Bars*>(sm + L::BARS);
    static_assert(sizeof(pg::Bars) <= 40 * sizeof(float), "barrier block");
    FOR_ITEMS(i, L::SMEM_FLOATS) sm[i] = 0.f;                     // operand padding (dead sample rows, k padding) must be finite
    TILE_SYNC();
    if (threadIdx.x == 0) pg::bars_init(bars, false);
    if ((threadIdx.x >> 5) == pg::kMmaWarp) pg::tmem_alloc2(tc::smem_u32(&bars->tmem_base));
    AdamScalars* adam_tab = reinterpret_cast<AdamScalars*>(sm + L::ADAMTAB);
    const SmallConsts C = stage_small_consts<S, L>(M, sm);
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
    tc::tc_fence_before();
    TILE_SYNC();
    tc::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pg::cluster_arrive();                                          // both CTAs' barriers exist before any remote arrive
    pg::cluster_wait();

    const pg::Consts K = {reinterpret_cast<const float4*>(M.pg_prior), reinterpret_cast<const float4*>(M.pg_fwd),
                          reinterpret_cast<const float4*>(M.pg_bwd)};
    const uint32_t sb = tc::smem_u32(sm);
    const pg::BOperands B = {sb + L::BPH * 4u, sb + L::BPL * 4u, sb + L::XT * 4u, sb + L::XLO * 4u, sb + L::QT * 4u, sb + L::LO * 4u};

    // ---- stage 1: global orientation + camera translation (camera_fitting_loss, smplify.py:70-91) ---------------------
    pair_forward<S, L>(M, C, K, sm, bars, B, rank, tmem_base, /*root_identity=*/true, /*do_select=*/false);
    ph_output_joints<S, L>(M, C, sm);
    TILE_SYNC();
    stage1_camera<S, L>(M, P, first, sm);
    TILE_SYNC();
    tile_zero_ignored_conf<S, L>(M, P, first, sm);
    zero_rows<S, L>(sm, L::ADM, 2 * kParams);
    TILE_SYNC();

    // ---- stage 2: body pose, betas, global orientation (body_fitting_loss, smplify.py:95-118) ---------------------------
    PAIR_CLK_DECL();
    for (int it = 0; it < P.num_iters; ++it) {
#if defined(PG_TRACE)
        if (threadIdx.x == 0 && blockIdx.x == 0) pg::g_pg_trace_on = (it == 7);
        if ((threadIdx.x >> 5) >= pg::kEpiWarp0 && (threadIdx.x >> 5) < pg::kMmaWarp && it == 7) { /* epilogue marks below */ }
        __syncthreads();
#endif
        pair_forward<S, L>(M, C, K, sm, bars, B, rank, tmem_base, false, true);
        PAIR_CLK_RESET();
        ph_output_joints<S, L>(M, C, sm);
        TILE_SYNC();
        if (threadIdx.x == 0) PAIR_CLK(4);
        ph_reprojection<S, L>(sm, P.focal, kSigma2, true);
        zero_rows<S, L>(sm, L::DG, 288);
        TILE_SYNC();
        if (threadIdx.x == 0) PAIR_CLK(5);
        if (P.loss_trace) {
            FOR_ITEMS(s, S) {
                const int b = first + s;
                float a = 0.f;
                for (int o = 0; o < kOut; ++o) a += sm[L::LOSSJ + o * S + s];
                a = ((a + sm[L::LOSSJ + 49 * S + s]) + sm[L::LOSSJ + 50 * S + s]) + sm[L::LOSSJ + 51 * S + s];
                if (b < P.batch) P.loss_trace[(size_t)(P.num_iters + it) * P.batch + b] = a;
            }
        }
        ph_joint_backward<S, L>(M, C, sm);
        TILE_SYNC();
        if (threadIdx.x == 0) PAIR_CLK(6);
        ph_pick_backward<S, L>(M, C, sm);
        if (threadIdx.x == 0) PAIR_CLK(7);
        pair_backward<S, L>(M, K, sm, bars, B, rank, tmem_base);
        PAIR_CLK_RESET();
        const AdamScalars sc = (it < kMaxIters) ? adam_tab[it] : adam_scalars(P, it);
        FOR_ITEMS(itj, kJoints * S) {
            const int s = itj % S, j = itj / S;
            float g[9], d[3];
            rotation_grad<S, L>(sm, j, s, g);
            rodrigues_bwd(sm[L::POSE + (3 * j + 0) * S + s], sm[L::POSE + (3 * j + 1) * S + s], sm[L::POSE + (3 * j + 2) * S + s], g, d);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int k = 3 * j + a;
                if (j > 0) d[a] += sm[L::GPR + (k - 3) * S + s];
                sm[L::POSE + k * S + s] = adam_update(sm[L::POSE + k * S + s], d[a], sm[L::ADM + k * S + s], sm[L::ADV + k * S + s], P.adam_c, sc);
            }
        }
        FOR_ITEMS(itb, kBetas * S) {
            const int s = itb % S, l = itb / S;
            const float beta = sm[L::BETA + l * S + s];
            const float g = beta_grad<S, L>(C, sm, l, s) + 2.f * kShapePriorW2 * beta;
            sm[L::BETA + l * S + s] = adam_update(beta, g, sm[L::ADM + (72 + l) * S + s], sm[L::ADV + (72 + l) * S + s], P.adam_c, sc);
        }
        TILE_SYNC();
        if (threadIdx.x == 0) PAIR_CLK(11);
    }

    // ---- final forward: joints, per-joint reprojection loss, operands of the vertex kernels ------------------------------
    pair_forward<S, L>(M, C, K, sm, bars, B, rank, tmem_base, false, false);
    ph_output_joints<S, L>(M, C, sm);
    TILE_SYNC();
    FOR_ITEMS(it, S * 147) {
        const int s = it / 147, k = it % 147, b = first + s;
        if (P.out_joints && b < P.batch) P.out_joints[(size_t)b * 147 + k] = sm[L::OUTJ + k * S + s];
    }
    tile_write_vertex_operands<S, L>(sm, first, P.batch, P.tc);
    TILE_SYNC();
    ph_reprojection<S, L>(sm, P.focal, kSigma2, false);
    TILE_SYNC();
    FOR_ITEMS(it, S * kOut) {
        const int s = it / kOut, o = it % kOut, b = first + s;
        if (b < P.batch) P.out_reproj[(size_t)b * kOut + o] = sm[L::LOSSJ + o * S + s];
    }
    FOR_ITEMS(it, S * 72) {
        const int s = it / 72, k = it % 72, b = first + s;
        if (P.out_pose && b < P.batch) P.out_pose[(size_t)b * 72 + k] = sm[L::POSE + k * S + s];
    }
    FOR_ITEMS(it, S * kBetas) {
        const int s = it / kBetas, k = it % kBetas, b = first + s;
        if (P.out_betas && b < P.batch) P.out_betas[(size_t)b * kBetas + k] = sm[L::BETA + k * S + s];
    }
    FOR_ITEMS(it, S * 3) {
        const int s = it / 3, k = it % 3, b = first + s;
        if (P.out_cam && b < P.batch) P.out_cam[(size_t)b * 3 + k] = sm[L::CAM + k * S + s];
    }
    if (P.out_packed) {
        constexpr int kPacked = 72 + kBetas + 3 + kOut;
        FOR_ITEMS(it, S * kPacked) {
            const int s = it / kPacked, k = it % kPacked, b = first + s;
            if (b >= P.batch) continue;
            float v;
            if (k < 72) v = sm[L::POSE + k * S + s];
            else if (k < 72 + kBetas) v = sm[L::BETA + (k - 72) * S + s];
            else if (k < 72 + kBetas + 3) v = sm[L::CAM + (k - 72 - kBetas) * S + s];
            else v = sm[L::LOSSJ + (k - 72 - kBetas - 3) * S + s];
            P.out_packed[(size_t)b * kPacked + k] = v;
        }
    }
    // ---- teardown: nobody may still address the peer's shared memory or TMEM ------------------------------------------------
    tc::tc_fence_before();
    pg::cluster_arrive();
    pg::cluster_wait();
    if ((threadIdx.x >> 5) == pg::kMmaWarp) { tc::tc_fence_after(); pg::tmem_dealloc2(tmem_base); }
}

#endif  // __CUDACC__

}  // namespace smplb200
