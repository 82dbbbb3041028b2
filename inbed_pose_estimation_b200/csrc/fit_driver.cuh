// Tile-level drivers built from the phases in fit_tile.cuh:
//   fit_tile            - SMPLify.__call__ (reference smplify/smplify.py:40-136) and get_fitting_loss (:138-172)
//   pose_forward_tile   - the per-sample half of SMPL.forward (models/smpl.py:21-33): joints, A, x
//   pose_backward_tile  - its gradient
// They are __host__ __device__ so the TEST-ONLY emulation build can run them on the host.
#pragma once
#include "fit_tile.cuh"
#include "lbs_tc.h"

namespace smplb200 {

// Optional per-phase cycle counters of the stage-2 loop (profiling builds only: -DSMPLB200_PHASE_CLOCKS;
// tools/phase_clocks.py).  Thread 0 of CTA 0 accumulates the clock64() deltas between phase marks.
#if defined(SMPLB200_PHASE_CLOCKS) && defined(__CUDACC__)
static __device__ unsigned long long g_phase_clocks[32];      // one copy per translation unit; kernels.cu owns the live one
#endif
#if defined(SMPLB200_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
#define PHASE_BEGIN() long long ph_t0 = clock64(); const bool ph_on = (threadIdx.x == 0 && blockIdx.x == 0)
#define PHASE_MARK(i) do { if (ph_on) { const long long t = clock64(); g_phase_clocks[i] += (unsigned long long)(t - ph_t0); ph_t0 = t; } } while (0)
#else
#define PHASE_BEGIN() ((void)0)
#define PHASE_MARK(i) ((void)0)
#endif

struct FitParams {
    int batch;
    int num_iters;            // per stage; 0 => only the final forward (get_fitting_loss)
    int zero_conf_first;      // get_fitting_loss zeroes the ignored confidences before anything else
    float focal;
    const float* init_pose;   // [B][72]
    const float* init_betas;  // [B][10]
    const float* init_cam;    // [B][3]
    const float* center;      // [B][2]
    float* keypoints;         // [B][49][3], confidences of the ignored joints are zeroed in place
    float* out_joints;        // [B][49][3]  (nullable)
    float* out_pose;          // [B][72]     (nullable)
    float* out_betas;         // [B][10]     (nullable)
    float* out_cam;           // [B][3]      (nullable)
    float* out_reproj;        // [B][49]
    float* out_packed;        // [B][134] (nullable) pose 72 | betas 10 | camera 3 | reprojection 49: the row a sharded refit gathers
    TcOperands tc;            // (nullable members) hi/lo tf32 operands of the tcgen05 vertex kernels for the final pose
    float* loss_trace;        // [2*num_iters][B] (nullable) per-sample loss of every iteration
    double lr, beta1, beta2;  // Adam hyper-parameters (smplify.py:79,107: lr=step_size, betas=(0.9, 0.999))
    AdamConsts adam_c;
};

// Bias-correction scalars of Adam step `it` (0-based), computed in float64 like torch does on the host.  The first
// kMaxIters steps of a stage are tabulated in shared memory once per tile; longer runs compute them on the fly.
SB_HD AdamScalars adam_scalars(const FitParams& P, int it) {
    const double bc1 = 1.0 - pow(P.beta1, (double)(it + 1));
    const double bc2 = 1.0 - pow(P.beta2, (double)(it + 1));
    AdamScalars sc;
    sc.step_size = (float)(P.lr / bc1);
    sc.bc2_sqrt = (float)sqrt(bc2);
    return sc;
}

constexpr float kSigma2 = 100.f * 100.f;              // gmof sigma (losses.py:28)
constexpr float kPosePriorW2 = (float)(4.78 * 4.78);  // pose_prior_weight ** 2
constexpr float kAnglePriorW2 = (float)(15.2 * 15.2); // angle_prior_weight ** 2
constexpr float kShapePriorW2 = 25.f;                 // shape_prior_weight ** 2
constexpr float kDepthW2 = 100.f * 100.f;             // depth_loss_weight ** 2 (losses.py:61)

// Stage the small constants in shared memory (device) or just point at them (host emulation).
// The caller must place a barrier before the first use.
template <int S, class L = TileLayout<S>>
SB_HD SmallConsts stage_small_consts(const ModelView& M, float* sm) {
    SmallConsts C;
#if defined(__CUDA_ARCH__)
    float* b = sm + L::CONSTS;
    float* JS = b; float* J0 = JS + 720; float* wkj = J0 + 72; float* Wp = wkj + 216;
    float* mu = Wp + 264; float* pmean = mu + 576; float* lognll = pmean + 576;
    FOR_ITEMS(i, 720) JS[i] = M.JS[i];
    FOR_ITEMS(i, 72) J0[i] = M.J0[i];
    FOR_ITEMS(i, 216) wkj[i] = M.wkj[i];
    FOR_ITEMS(i, 264) Wp[i] = M.Wp[i];
    FOR_ITEMS(i, 576) { mu[i] = M.gmm_means[i]; pmean[i] = M.gmm_pmean[i]; }
    FOR_ITEMS(i, 8) lognll[i] = M.gmm_lognll[i];
    C.JS = JS; C.J0 = J0; C.wkj = wkj; C.Wp = Wp; C.mu = mu; C.pmean = pmean; C.lognll = lognll;
#else
    (void)sm;
    C.JS = M.JS; C.J0 = M.J0; C.wkj = M.wkj; C.Wp = M.Wp; C.mu = M.gmm_means; C.pmean = M.gmm_pmean; C.lognll = M.gmm_lognll;
#endif
    return C;
}

// Kinematic chain + folded forward GEMM (independent of each other: the chain reads the rotations and rest joints, the
// GEMM reads x).  On the device with the standard 384-thread tile the chain's level sweeps run on warp 11, which no GEMM
// work item lands on, concurrently with the GEMM; otherwise one after the other.  Ends with a tile barrier.
template <int S, class L = TileLayout<S>>
SB_HD void ph_chain_and_gemm_forward(const ModelView& M, float* sm) {
#if defined(__CUDA_ARCH__)
    if (kFastGemm<S> && TILE_NT == kFitTileThreads) {
        if (TILE_TID >= kChainWarpFirstThread) ph_chain_forward<S, L>(M, sm, Grp{TILE_TID - kChainWarpFirstThread, 32, 1});
        else ph_fold_gemm_forward<S, L>(M, sm);
        TILE_SYNC();
        return;
    }
#endif
    ph_chain_forward<S, L>(M, sm, grp_tile());           // ends with a barrier
    ph_fold_gemm_forward<S, L>(M, sm);
    TILE_SYNC();
}

// Folded backward GEMM + reverse chain sweep, overlapped the same way (the sweep needs dL/dG from ph_joint_backward only;
// the GEMM's dL/dx is added to dL/dR afterwards in rotation_grad).  Ends with a tile barrier.
template <int S, class L = TileLayout<S>>
SB_HD void ph_gemm_and_chain_backward(const ModelView& M, float* sm) {
#if defined(__CUDA_ARCH__)
    if (kFastGemm<S> && TILE_NT == kFitTileThreads) {
        if (TILE_TID >= kChainWarpFirstThread) ph_chain_backward<S, L>(M, sm, Grp{TILE_TID - kChainWarpFirstThread, 32, 1});
        ph_fold_gemm_backward<S, L>(M, sm);              // warp 11 owns no GEMM range: it only joins the barriers and the combine
        TILE_SYNC();
        return;
    }
#endif
    ph_fold_gemm_backward<S, L>(M, sm);
    TILE_SYNC();
    ph_chain_backward<S, L>(M, sm, grp_tile());
    TILE_SYNC();
}

// forward through the folded joint model for the tile's current parameters
template <int S, class L = TileLayout<S>>
SB_HD void tile_forward(const ModelView& M, const SmallConsts& C, float* sm, bool from_axis_angle, bool root_identity) {
    ph_pose_features<S, L>(sm, from_axis_angle, root_identity);
    ph_rest_joints<S, L>(C, sm);
    TILE_SYNC();
    ph_chain_and_gemm_forward<S, L>(M, sm);
    ph_output_joints<S, L>(M, C, sm);
    TILE_SYNC();
}

// Skinning transforms A = [G^R | A^t] and blend coefficients x of the tile's current pose, written as the hi/lo tf32
// split operands of the tcgen05 vertex kernels (lbs_tc.cu): x [B][224], transforms [B][12 entries][24 joints + 8 pad].
template <int S, class L = TileLayout<S>>
SB_HD void tile_write_vertex_operands(float* sm, int first, int batch, const TcOperands& tc) {
    if (tc.x_hi) {
        FOR_ITEMS(it, S * kXPad) {
            const int s = it / kXPad, k = it % kXPad, b = first + s;
            if (b >= batch) continue;
            const float x = sm[L::x(k, s)];
            const float hi = tf32_round(x);
            tc.x_hi[(size_t)b * kXPad + k] = hi;
            tc.x_lo[(size_t)b * kXPad + k] = x - hi;
        }
    }
    if (tc.ae_hi) {
        FOR_ITEMS(it, S * kAeRow) {
            const int s = it / kAeRow, r = it % kAeRow, e = r / 32, j = r % 32, b = first + s;
            if (b >= batch) continue;
            float a = 0.f;
            if (j < kJoints) a = (e % 4 == 3) ? sm[L::AT + (3 * j + e / 4) * S + s] : sm[L::GW + (j * 12 + e) * S + s];
            const float hi = tf32_round(a);
            const size_t o = (size_t)b * kAeRow + r;
            tc.ae_hi[o] = hi;
            tc.ae_lo[o] = a - hi;
        }
    }
}

// Stage 1 (camera_fitting_loss, smplify.py:70-91): only the root rotation and the camera move, so all
// 49 joints are an affine image of the root-identity pose:  joint = R0 (rest - sigma J0) + sigma J0.
// One thread per sample runs the whole stage in registers.
template <int S, class L = TileLayout<S>>
SB_HD void stage1_camera(const ModelView& M, const FitParams& P, int first, float* sm) {
    const AdamScalars* adam_tab = reinterpret_cast<const AdamScalars*>(sm + L::ADAMTAB);
    FOR_ITEMS(s, S) {
        const int b = first + s;
        float th[3], t[3], mm[6], vv[6];
#pragma unroll
        for (int a = 0; a < 3; ++a) { th[a] = sm[L::POSE + a * S + s]; t[a] = sm[L::CAM + a * S + s]; }
#pragma unroll
        for (int a = 0; a < 6; ++a) { mm[a] = 0.f; vv[a] = 0.f; }
        const float tz0 = t[2];
        const float cx = sm[L::CEN + 0 * S + s], cy = sm[L::CEN + 1 * S + s];
        float cmin = sm[L::KP + (3 * M.cam_op[0] + 2) * S + s];
#pragma unroll
        for (int q = 1; q < 4; ++q) cmin = fminf(cmin, sm[L::KP + (3 * M.cam_op[q] + 2) * S + s]);
        const bool use_op = cmin > 0.f;                       // losses.py:83
        const float J0[3] = {sm[L::JR + 0 * S + s], sm[L::JR + 1 * S + s], sm[L::JR + 2 * S + s]};
        float av[4][3], bv[4][3], kx[4], ky[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int o = use_op ? M.cam_op[q] : M.cam_gt[q];
            const float sig = M.sigma_src[M.joint_map[o]];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                bv[q][c] = sig * J0[c];
                av[q][c] = sm[L::OUTJ + (3 * o + c) * S + s] - bv[q][c];
            }
            kx[q] = sm[L::KP + (3 * o + 0) * S + s];
            ky[q] = sm[L::KP + (3 * o + 1) * S + s];
        }
        for (int it = 0; it < P.num_iters; ++it) {
            float R[9], dR[9], dt[3] = {0.f, 0.f, 0.f};
            rodrigues_fwd(th[0], th[1], th[2], R);
#pragma unroll
            for (int e = 0; e < 9; ++e) dR[e] = 0.f;
            float loss = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float Pw[3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    Pw[r] = ((R[r * 3 + 0] * av[q][0] + R[r * 3 + 1] * av[q][1] + R[r * 3 + 2] * av[q][2]) + bv[q][r]) + t[r];
                const float px = Pw[0] / Pw[2], py = Pw[1] / Pw[2];
                const float ex = kx[q] - (P.focal * px + cx), ey = ky[q] - (P.focal * py + cy);
                loss += ex * ex + ey * ey;
                const float gu = -2.f * ex * P.focal, gv = -2.f * ey * P.focal;
                const float dP[3] = {gu / Pw[2], gv / Pw[2], -(gu * px + gv * py) / Pw[2]};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    dR[r * 3 + 0] += dP[r] * av[q][0]; dR[r * 3 + 1] += dP[r] * av[q][1]; dR[r * 3 + 2] += dP[r] * av[q][2];
                    dt[r] += dP[r];
                }
            }
            const float dz = t[2] - tz0;
            loss += kDepthW2 * (dz * dz);
            dt[2] += 2.f * kDepthW2 * dz;
            if (P.loss_trace && b < P.batch) P.loss_trace[(size_t)it * P.batch + b] = loss;
            float dth[3];
            rodrigues_bwd(th[0], th[1], th[2], dR, dth);
            const AdamScalars sc = (it < kMaxIters) ? adam_tab[it] : adam_scalars(P, it);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                th[a] = adam_update(th[a], dth[a], mm[a], vv[a], P.adam_c, sc);
                t[a] = adam_update(t[a], dt[a], mm[3 + a], vv[3 + a], P.adam_c, sc);
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { sm[L::POSE + a * S + s] = th[a]; sm[L::CAM + a * S + s] = t[a]; }
    }
}

template <int S, class L = TileLayout<S>>
SB_HD void tile_zero_ignored_conf(const ModelView& M, const FitParams& P, int first, float* sm) {
    FOR_ITEMS(it, M.num_ign * S) {
        const int s = it % S, o = M.ign_joints[it / S], b = first + s;
        sm[L::KP + (3 * o + 2) * S + s] = 0.f;
        if (b < P.batch) P.keypoints[((size_t)b * kOut + o) * 3 + 2] = 0.f;   // in place, smplify.py:105 / :156
    }
}

// One Adam step of stage 2 on the tile's 72 pose entries and 10 betas: chain rule through the rotations (dL/dR from the chain
// sweep + dL/dx of the folded GEMM, in XT) and the shape features, plus the prior gradients.
template <int S, class L = TileLayout<S>>
SB_HD void tile_adam_step(const SmallConsts& C, const FitParams& P, float* sm, const AdamScalars& sc) {
    FOR_ITEMS(itj, kJoints * S) {
        const int s = itj % S, j = itj / S;
        float g[9], d[3];
        rotation_grad<S, L>(sm, j, s, g);
        rodrigues_bwd(sm[L::POSE + (3 * j + 0) * S + s], sm[L::POSE + (3 * j + 1) * S + s],
                      sm[L::POSE + (3 * j + 2) * S + s], g, d);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int k = 3 * j + a;
            if (j > 0) d[a] += sm[L::GPR + (k - 3) * S + s];
            sm[L::POSE + k * S + s] = adam_update(sm[L::POSE + k * S + s], d[a], sm[L::ADM + k * S + s],
                                                  sm[L::ADV + k * S + s], P.adam_c, sc);
        }
    }
    FOR_ITEMS(itb, kBetas * S) {
        const int s = itb % S, l = itb / S;
        const float beta = sm[L::BETA + l * S + s];
        const float g = beta_grad<S, L>(C, sm, l, s) + 2.f * kShapePriorW2 * beta;
        sm[L::BETA + l * S + s] = adam_update(beta, g, sm[L::ADM + (72 + l) * S + s], sm[L::ADV + (72 + l) * S + s],
                                              P.adam_c, sc);
    }
}

// Results of a tile whose final forward has been run: joints, operands of the vertex kernels, per-joint reprojection loss,
// parameters, packed rows.  Contains tile barriers: the whole tile calls it.
template <int S, class L = TileLayout<S>>
SB_HD void tile_write_outputs(const FitParams& P, int first, float* sm) {
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
}

template <int S, class L = TileLayout<S>>
SB_HD void fit_tile(const ModelView& M, const FitParams& P, int first, float* sm) {
    // per-iteration Adam scalars, computed in float64 like torch does on the host
    AdamScalars* adam_tab = reinterpret_cast<AdamScalars*>(sm + L::ADAMTAB);
    const SmallConsts C = stage_small_consts<S, L>(M, sm);
    FOR_ITEMS(t, (P.num_iters < kMaxIters ? P.num_iters : kMaxIters)) adam_tab[t] = adam_scalars(P, t);
    // ---- load the tile ---------------------------------------------------------------------
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
    TILE_SYNC();
    if (P.zero_conf_first) { tile_zero_ignored_conf<S, L>(M, P, first, sm); TILE_SYNC(); }

    if (P.num_iters > 0) {
        // ---- stage 1: global orientation + camera translation --------------------------------
        tile_forward<S, L>(M, C, sm, true, /*root_identity=*/true);
        stage1_camera<S, L>(M, P, first, sm);
        TILE_SYNC();
        tile_zero_ignored_conf<S, L>(M, P, first, sm);
        zero_rows<S, L>(sm, L::ADM, 2 * kParams);
        TILE_SYNC();

        // ---- stage 2: body pose, betas, global orientation (body_fitting_loss) ----------------
        PHASE_BEGIN();
        for (int it = 0; it < P.num_iters; ++it) {
            ph_prior_quadratic<S, L>(M, C, sm);
            TILE_SYNC();
            PHASE_MARK(0);
            ph_prior_select<S, L>(M, C, sm, kPosePriorW2, kAnglePriorW2, kShapePriorW2);
            TILE_SYNC();
            PHASE_MARK(1);
            ph_pose_features<S, L>(sm, true, false);
            ph_rest_joints<S, L>(C, sm);
            TILE_SYNC();
            PHASE_MARK(2);
            PHASE_MARK(3);
            ph_chain_and_gemm_forward<S, L>(M, sm);
            PHASE_MARK(4);
            ph_output_joints<S, L>(M, C, sm);
            TILE_SYNC();
            PHASE_MARK(5);
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
            PHASE_MARK(6);
            ph_joint_backward<S, L>(M, C, sm);
            TILE_SYNC();
            PHASE_MARK(7);
            ph_pick_backward<S, L>(M, C, sm);
            TILE_SYNC();
            PHASE_MARK(8);
            PHASE_MARK(9);
            ph_gemm_and_chain_backward<S, L>(M, sm);
            PHASE_MARK(10);
            const AdamScalars sc = (it < kMaxIters) ? adam_tab[it] : adam_scalars(P, it);
            tile_adam_step<S, L>(C, P, sm, sc);
            TILE_SYNC();
            PHASE_MARK(11);
        }
    }

    // ---- final forward: joints, per-joint reprojection loss, A and x for the vertex kernel --------
    tile_forward<S, L>(M, C, sm, true, false);
    tile_write_outputs<S, L>(P, first, sm);
}

// --------------------------------------------------------------------------------------------------
// The three prior terms of body_fitting_loss on their own (smplify/losses.py:46-52: 4.78^2 * MaxMixturePrior
// (smplify/prior.py:181-196) + 15.2^2 * angle_prior (losses.py:19-24) + 5^2 * |betas|^2) with their gradients, through the
// SAME tile phases the fit kernel runs every iteration - the isolated check of rows a15, a16 (prior part), a19.
// --------------------------------------------------------------------------------------------------
struct PriorParams {
    int batch;
    const float* pose;        // [B][72] full pose (global orientation first; the prior reads entries 3..71)
    const float* betas;       // [B][10]
    float* terms;             // [B][3]  weighted pose prior, angle prior, shape prior
    float* components;        // [B][8]  0.5 d^T P d - log(nll_weight) of every mixture component (nullable)
    int* argmin;              // [B]     selected component (nullable)
    float* grad_body_pose;    // [B][69] d(sum of the three terms)/d body_pose (nullable)
    float* grad_betas;        // [B][10] (nullable)
};

template <int S, class L = TileLayout<S>>
SB_HD void prior_tile(const ModelView& M, const PriorParams& P, int first, float* sm) {
    const SmallConsts C = stage_small_consts<S, L>(M, sm);
    FOR_ITEMS(it, S * 72) {
        const int s = it / 72, k = it % 72, b = first + s;
        sm[L::POSE + k * S + s] = (b < P.batch) ? P.pose[(size_t)b * 72 + k] : 0.f;
    }
    FOR_ITEMS(it, S * kBetas) {
        const int s = it / kBetas, k = it % kBetas, b = first + s;
        sm[L::BETA + k * S + s] = (b < P.batch) ? P.betas[(size_t)b * kBetas + k] : 0.f;
    }
    TILE_SYNC();
    ph_prior_quadratic<S, L>(M, C, sm);
    TILE_SYNC();
    ph_prior_select<S, L>(M, C, sm, kPosePriorW2, kAnglePriorW2, kShapePriorW2);
    TILE_SYNC();
    FOR_ITEMS(it, S * 3) {
        const int s = it / 3, k = it % 3, b = first + s;
        if (b < P.batch) P.terms[(size_t)b * 3 + k] = sm[L::LOSSJ + (49 + k) * S + s];
    }
    FOR_ITEMS(it, S * kGauss) {
        const int s = it / kGauss, g = it % kGauss, b = first + s;
        if (P.components && b < P.batch) P.components[(size_t)b * kGauss + g] = sm[L::MISC + g * S + s];
    }
    FOR_ITEMS(s, S) {
        const int b = first + s;
        if (P.argmin && b < P.batch) P.argmin[b] = (int)sm[L::MISC + 8 * S + s];
    }
    FOR_ITEMS(it, S * kPriorDim) {
        const int s = it / kPriorDim, i = it % kPriorDim, b = first + s;
        if (P.grad_body_pose && b < P.batch) P.grad_body_pose[(size_t)b * kPriorDim + i] = sm[L::GPR + i * S + s];
    }
    FOR_ITEMS(it, S * kBetas) {
        const int s = it / kBetas, l = it % kBetas, b = first + s;
        if (P.grad_betas && b < P.batch) P.grad_betas[(size_t)b * kBetas + l] = 2.f * kShapePriorW2 * sm[L::BETA + l * S + s];
    }
}

// --------------------------------------------------------------------------------------------------
// SMPL.forward / backward, per-sample half
// --------------------------------------------------------------------------------------------------
struct PoseParams {
    int batch;
    int rotmat_mode;          // 0: pose is axis-angle [B][72]; 1: rotation matrices [B][24][9]
    const float* pose;
    const float* betas;       // [B][10]
    float* joints;            // [B][49][3]
    TcOperands tc;            // (nullable members) hi/lo tf32 operands of the tcgen05 vertex kernels
    // backward only
    const float* d_joints;    // [B][49][3] (nullable)
    const float* dA_part;     // [nsplit_a][B][12][24] (nullable) partial dL/dA of the dA kernel, entry-major
    const float* dx_part;     // [nsplit_x][B][224]    (nullable) partial dL/dx of the dx GEMM
    int nsplit_a, nsplit_x;
    float* d_pose;            // [B][72] or [B][24][9]
    float* d_betas;           // [B][10]
};

// p[0] + p[stride] + ... (n terms of four floats, in that order; p 16-byte aligned, stride a multiple of 4).  Eight loads are
// issued before the first add: the partials come from HBM and a dependent load-add chain would pay the full memory latency
// per split.
SB_HD float4 sum_partials4(const float* p, size_t stride, int n) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i0 = 0; i0 < n; i0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            v[u] = (i0 + u < n) ? *reinterpret_cast<const float4*>(p + (size_t)(i0 + u) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    return a;
}

template <int S, class L = TileLayout<S>>
SB_HD void pose_load(const PoseParams& P, int first, float* sm) {
    if (P.rotmat_mode) {
        FOR_ITEMS(it, S * 216) {
            const int s = it / 216, k = it % 216, b = first + s;
            sm[L::RM + k * S + s] = (b < P.batch) ? P.pose[(size_t)b * 216 + k] : ((k % 9) % 4 == 0 ? 1.f : 0.f);
        }
    } else {
        FOR_ITEMS(it, S * 72) {
            const int s = it / 72, k = it % 72, b = first + s;
            sm[L::POSE + k * S + s] = (b < P.batch) ? P.pose[(size_t)b * 72 + k] : 0.f;
        }
    }
    FOR_ITEMS(it, S * kBetas) {
        const int s = it / kBetas, k = it % kBetas, b = first + s;
        sm[L::BETA + k * S + s] = (b < P.batch) ? P.betas[(size_t)b * kBetas + k] : 0.f;
    }
    TILE_SYNC();
}

template <int S, class L = TileLayout<S>>
SB_HD void pose_forward_tile(const ModelView& M, const PoseParams& P, int first, float* sm) {
    const SmallConsts C = stage_small_consts<S, L>(M, sm);
    pose_load<S, L>(P, first, sm);
    tile_forward<S, L>(M, C, sm, !P.rotmat_mode, false);
    FOR_ITEMS(it, S * 147) {
        const int s = it / 147, k = it % 147, b = first + s;
        if (P.joints && b < P.batch) P.joints[(size_t)b * 147 + k] = sm[L::OUTJ + k * S + s];
    }
    tile_write_vertex_operands<S, L>(sm, first, P.batch, P.tc);
}

template <int S, class L = TileLayout<S>>
SB_HD void pose_backward_tile(const ModelView& M, const PoseParams& P, int first, float* sm) {
    const SmallConsts C = stage_small_consts<S, L>(M, sm);
    pose_load<S, L>(P, first, sm);
    tile_forward<S, L>(M, C, sm, !P.rotmat_mode, false);
    FOR_ITEMS(it, S * 147) {
        const int s = it / 147, k = it % 147, b = first + s;
        sm[L::OUTJ + k * S + s] = (P.d_joints && b < P.batch) ? P.d_joints[(size_t)b * 147 + k] : 0.f;
    }
    FOR_ITEMS(it, S * 72) {
        // four consecutive entries of the partials' own (entry-major [12][24]) order per item: 16-byte loads, a warp reads
        // 512 contiguous bytes per split, eight splits in flight
        const int s = it / 72, k4 = it % 72, b = first + s;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.dA_part && b < P.batch) a = sum_partials4(P.dA_part + (size_t)b * 288 + 4 * k4, (size_t)P.batch * 288, P.nsplit_a);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * k4 + u, e = k / kJoints, j = k % kJoints;
            sm[L::DG + (j * 12 + e) * S + s] = av[u];
        }
    }
    TILE_SYNC();
    ph_joint_backward<S, L>(M, C, sm);
    TILE_SYNC();
    ph_pick_backward<S, L>(M, C, sm);
    TILE_SYNC();
    ph_gemm_and_chain_backward<S, L>(M, sm);
    if (P.dx_part) {
        FOR_ITEMS(it, S * (kXPad / 4)) {
            const int s = it / (kXPad / 4), k4 = it % (kXPad / 4), b = first + s;
            if (b < P.batch) {
                const float4 a = sum_partials4(P.dx_part + (size_t)b * kXPad + 4 * k4, (size_t)P.batch * kXPad, P.nsplit_x);
                sm[L::XT + (4 * k4 + 0) * S + s] += a.x;
                sm[L::XT + (4 * k4 + 1) * S + s] += a.y;
                sm[L::XT + (4 * k4 + 2) * S + s] += a.z;
                sm[L::XT + (4 * k4 + 3) * S + s] += a.w;
            }
        }
        TILE_SYNC();
    }
    FOR_ITEMS(it, kJoints * S) {
        const int s = it % S, j = it / S, b = first + s;
        float g[9];
        rotation_grad<S, L>(sm, j, s, g);
        if (b >= P.batch) continue;
        if (P.rotmat_mode) {
#pragma unroll
            for (int e = 0; e < 9; ++e) P.d_pose[(size_t)b * 216 + j * 9 + e] = g[e];
        } else {
            float d[3];
            rodrigues_bwd(sm[L::POSE + (3 * j + 0) * S + s], sm[L::POSE + (3 * j + 1) * S + s],
                          sm[L::POSE + (3 * j + 2) * S + s], g, d);
#pragma unroll
            for (int a = 0; a < 3; ++a) P.d_pose[(size_t)b * 72 + 3 * j + a] = d[a];
        }
    }
    FOR_ITEMS(it, kBetas * S) {
        const int s = it % S, l = it / S, b = first + s;
        const float g = beta_grad<S, L>(C, sm, l, s);
        if (b < P.batch) P.d_betas[(size_t)b * kBetas + l] = g;
    }
}

}  // namespace smplb200
