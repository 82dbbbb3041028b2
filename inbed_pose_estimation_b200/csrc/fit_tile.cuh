// Per-tile (S samples per CTA) phases of the folded SMPL joint model, its backward pass,
// the SMPLify losses and the Adam update.  Every phase is a parallel-for over "items"
// followed by a block barrier; all cross-phase state lives in the tile's shared memory
// ([index][S] layout, sample fastest, so lanes of a warp touch consecutive words).
//
// Reference semantics restated here (paths into /root/reference, smplx rows per SURVEY.md §8a):
//   a8  smplx batch_rodrigues          -> rodrigues_fwd / rodrigues_bwd
//   a9  pose_feature = R[1:] - I       -> ph_pose_features
//   a7  J = J_regressor . v_shaped     -> ph_rest_joints   (folded: J0 + JS.beta)
//   a10 batch_rigid_transform          -> ph_chain_forward / ph_chain_backward
//   a11/a12/a5 skinning + joint regressors + joint_map -> folded GEMM (Cf) + ph_joint_*
//   a13 perspective_projection, a14 gmof, a16 body_fitting_loss (smplify/losses.py:26-58)
//   a17 camera_fitting_loss (losses.py:60-90), a19 MaxMixturePrior (prior.py:181-196)
//   a15 angle_prior (losses.py:19-24), a20 torch.optim.Adam (_single_tensor_adam)
#pragma once
#include <math.h>
#include "smpl_common.h"

namespace smplb200 {

#if defined(__CUDA_ARCH__)
#define TILE_TID ((int)threadIdx.x)
#define TILE_NT ((int)blockDim.x)
#define TILE_SYNC() __syncthreads()
#else
#define TILE_TID 0
#define TILE_NT 1
#define TILE_SYNC() ((void)0)
#endif
#define FOR_ITEMS(it, n) for (int it = TILE_TID; it < (n); it += TILE_NT)
#define TILE_SYNC_NONE() ((void)0)

// The threads that execute a phase: the whole tile (block barrier between sub-steps) or a single warp (warp barrier).
// The latency-bound kinematic-chain sweeps run on the one warp the folded GEMMs leave idle, concurrently with them.
// warp_only: 0 = the whole tile (block barrier), 1 = one warp, 2 = a group of whole warps meeting on named barrier 2
struct Grp { int tid, nt, warp_only; };
SB_HD Grp grp_tile() { return Grp{TILE_TID, TILE_NT, 0}; }
SB_HD void grp_sync(const Grp& g) {
#if defined(__CUDA_ARCH__)
    if (g.warp_only == 1) __syncwarp();
    else if (g.warp_only == 2) asm volatile("bar.sync 2, %0;" ::"r"(g.nt) : "memory");
    else if (g.warp_only == 3) asm volatile("bar.sync 1, %0;" ::"r"(g.nt) : "memory");      // a second named group beside a `2` group
    else __syncthreads();
#else
    (void)g;
#endif
}
#define FOR_ITEMS_G(it, n, g) for (int it = (g).tid; it < (n); it += (g).nt)
constexpr int kChainWarpFirstThread = 352;      // warp 11 of a 384-thread tile: no work item of either folded GEMM lands on it

// Small per-model constants every iteration touches; the kernels stage them in shared memory once
// (global / L2 latency would otherwise be exposed in the short per-sample phases).
struct SmallConsts {
    const float* JS;      // [72][10]
    const float* J0;      // [72]
    const float* wkj;     // [9][24]
    const float* Wp;      // [11][24]
    const float* mu;      // [8][72]  gmm means (padded)
    const float* pmean;   // [8][72]  Psym . mean (padded)
    const float* lognll;  // [8]
};
constexpr int kSmallConstFloats = 720 + 72 + 216 + 264 + 576 + 576 + 8;   // 2432

SB_HD float2 ld_const2(const float2* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

template <int S_>
struct TileLayout {
    static constexpr int S = S_;
    static constexpr int LDQ = S + 4;               // padded row of QT: conflict-free 16-byte column stores
    static constexpr int POSE = 0;                  // [72][S]
    static constexpr int BETA = POSE + 72 * S;      // [10][S]
    static constexpr int CAM = BETA + 10 * S;       // [3][S]
    static constexpr int CEN = CAM + 3 * S;         // [2][S]
    static constexpr int KP = CEN + 2 * S;          // [147][S]   (x, y, conf) per output joint
    static constexpr int RM = KP + 147 * S;         // [216][S]   rotations; reused for dL/dR
    static constexpr int XT = RM + 216 * S;         // [224][S]   x = [1, beta, feat]; reused for dL/dx
    static constexpr int JR = XT + kXPad * S;       // [72][S]    rest joints
    static constexpr int GW = JR + 72 * S;          // [288][S]   world transforms (3x4 row-major)
    static constexpr int AT = GW + 288 * S;         // [72][S]    A_j translation column
    static constexpr int QT = AT + 72 * S;          // [704][LDQ] folded GEMM output; reused for dL/dQ and prior scratch
    static constexpr int OUTJ = QT + kQPad * LDQ;   // [147][S]   output joints; reused for dL/djoints
    static constexpr int DG = OUTJ + 147 * S;       // [288][S]
    static constexpr int DJ = DG + 288 * S;         // [72][S]
    static constexpr int GPR = DJ + 72 * S;         // [69][S]    prior gradient w.r.t. body pose
    static constexpr int LOSSJ = GPR + 69 * S;      // [52][S]    49 reprojection terms, gmm, angle, shape
    static constexpr int ADM = LOSSJ + 52 * S;      // [82][S]
    static constexpr int ADV = ADM + 82 * S;        // [82][S]
    static constexpr int MISC = ADV + 82 * S;       // [16][S]
    static constexpr int TOTAL = MISC + 16 * S;
    static constexpr int ADAMTAB = TOTAL;                       // [kMaxIters] AdamScalars
    static constexpr int CONSTS = ADAMTAB + 2 * kMaxIters;      // SmallConsts arrays
    static constexpr int SMEM_FLOATS = CONSTS + kSmallConstFloats;
    static_assert(kGauss * kPriorPad * S <= kQPad * LDQ, "prior scratch must fit in the QT region");
    static_assert(2 * kXPad * S <= kQPad * LDQ, "backward GEMM partials must fit in the QT region");
    // index functions of the two arrays whose layout the GEMM implementation dictates (the pair kernel keeps them as
    // tensor-core operands, fit_pair.cuh): Q / dQ [n][sample] and x [m][sample]
    static constexpr bool kPair = false;
    SB_HD static int q(int n, int s) { return QT + n * LDQ + s; }
    SB_HD static int x(int m, int s) { return XT + m * S_ + s; }
    SB_HD static int pd(int g, int i, int s) { return QT + (g * kPriorPad + i) * S_ + s; }      // prior scratch Psym (bp - mean)
    SB_HD static void store_dq(float* sm, int n, int s, float v) { sm[q(n, s)] = v; }
    SB_HD static void store_x(float* sm, int m, int s, float v) { sm[x(m, s)] = v; }
};

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
// smplx.lbs.batch_rodrigues for one joint: angle = |r + 1e-8|, n = r / angle,
// R = (I + sin K) + (1 - cos) K K.
SB_HD void rodrigues_fwd(float rx, float ry, float rz, float* R) {
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float ang = sqrtf(ax * ax + ay * ay + az * az);
    const float nx = rx / ang, ny = ry / ang, nz = rz / ang;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    const float oc = 1.0f - cs;
    // K = [[0,-nz,ny],[nz,0,-nx],[-ny,nx,0]];  KK = K*K
    const float kk00 = -nz * nz - ny * ny, kk01 = nx * ny, kk02 = nx * nz;
    const float kk11 = -nz * nz - nx * nx, kk12 = ny * nz, kk22 = -ny * ny - nx * nx;
    R[0] = 1.0f + oc * kk00;
    R[1] = (-sn * nz) + oc * kk01;
    R[2] = (sn * ny) + oc * kk02;
    R[3] = (sn * nz) + oc * kk01;
    R[4] = 1.0f + oc * kk11;
    R[5] = (-sn * nx) + oc * kk12;
    R[6] = (-sn * ny) + oc * kk02;
    R[7] = (sn * nx) + oc * kk12;
    R[8] = 1.0f + oc * kk22;
}

// Gradient of the map above: g = dL/dR (row-major 3x3) -> dL/dr.
SB_HD void rodrigues_bwd(float rx, float ry, float rz, const float* g, float* dr) {
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float ang = sqrtf(ax * ax + ay * ay + az * az);
    const float inv = 1.0f / ang;
    const float nx = rx / ang, ny = ry / ang, nz = rz / ang;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    const float oc = 1.0f - cs;
    const float K[9] = {0.f, -nz, ny, nz, 0.f, -nx, -ny, nx, 0.f};
    const float KK[9] = {-nz * nz - ny * ny, nx * ny, nx * nz,
                         nx * ny, -nz * nz - nx * nx, ny * nz,
                         nx * nz, ny * nz, -ny * ny - nx * nx};
    float d_sin = 0.f, d_oc = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        d_sin += g[e] * K[e];
        d_oc += g[e] * KK[e];
    }
    // dK = sin * g + (1-cos) * (g K^T + K^T g)
    float dK[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) a += g[r * 3 + k] * K[c * 3 + k] + K[k * 3 + r] * g[k * 3 + c];
            dK[r * 3 + c] = sn * g[r * 3 + c] + oc * a;
        }
    const float dnx = dK[7] - dK[5], dny = dK[2] - dK[6], dnz = dK[3] - dK[1];
    float d_ang = d_sin * cs + d_oc * sn;
    d_ang -= (dnx * rx + dny * ry + dnz * rz) * inv * inv;
    dr[0] = dnx * inv + d_ang * ax * inv;
    dr[1] = dny * inv + d_ang * ay * inv;
    dr[2] = dnz * inv + d_ang * az * inv;
}

// Geman-McClure (losses.py:11-17) and its derivative.
SB_HD float gmof_val(float x, float sigma2) { const float x2 = x * x; return (sigma2 * x2) / (sigma2 + x2); }
SB_HD float gmof_grad(float x, float sigma2) { const float d = sigma2 + x * x; return (2.0f * sigma2 * sigma2 * x) / (d * d); }

struct AdamConsts { float lerp_w, beta2, w2, eps; };

// torch.optim.Adam single-tensor step, fp32 (adam.py _single_tensor_adam).
SB_HD float adam_update(float p, float g, float& m, float& v, const AdamConsts& c, const AdamScalars& t) {
    m = m + c.lerp_w * (g - m);
    v = v * c.beta2;
    v = v + (c.w2 * g) * g;
    const float denom = sqrtf(v) / t.bc2_sqrt + c.eps;
    return p + (-t.step_size) * (m / denom);
}

// ------------------------------------------------------------------------------------------------
// forward phases
// ------------------------------------------------------------------------------------------------
// Rotations of all joints from axis-angle + pose features / betas into x.  root_identity forces
// R_0 = I (stage-1 hoisting: the rest pose every root rotation is applied to).
template <int S, class L = TileLayout<S>>
SB_HD void ph_pose_features(float* sm, bool from_axis_angle, bool root_identity) {
    FOR_ITEMS(it, kJoints * S) {
        const int s = it % S, j = it / S;
        float R[9];
        if (from_axis_angle) {
            rodrigues_fwd(sm[L::POSE + (3 * j + 0) * S + s], sm[L::POSE + (3 * j + 1) * S + s],
                          sm[L::POSE + (3 * j + 2) * S + s], R);
        } else {
#pragma unroll
            for (int e = 0; e < 9; ++e) R[e] = sm[L::RM + (j * 9 + e) * S + s];
        }
        if (j == 0 && root_identity) {
#pragma unroll
            for (int e = 0; e < 9; ++e) R[e] = (e % 4 == 0) ? 1.f : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 9; ++e) {
            sm[L::RM + (j * 9 + e) * S + s] = R[e];
            if (j > 0) L::store_x(sm, 11 + (j - 1) * 9 + e, s, R[e] - ((e % 4 == 0) ? 1.f : 0.f));
        }
    }
    FOR_ITEMS(it, (1 + kBetas + (kXPad - kX)) * S) {
        const int s = it % S, r = it / S;
        if (r == 0) L::store_x(sm, 0, s, 1.f);
        else if (r <= kBetas) L::store_x(sm, r, s, sm[L::BETA + (r - 1) * S + s]);
        else L::store_x(sm, kX + r - 1 - kBetas, s, 0.f);
    }
}

// J = J0 + JS.beta  (== J_regressor.(v_template + shapedirs.beta), folded in float64 at model creation)
template <int S, class L = TileLayout<S>>
SB_HD void ph_rest_joints(const SmallConsts& C, float* sm) {
    FOR_ITEMS(it, 72 * S) {
        const int s = it % S, jc = it / S;
        float a = C.J0[jc];
#pragma unroll
        for (int l = 0; l < kBetas; ++l) a += C.JS[jc * kBetas + l] * sm[L::BETA + l * S + s];
        sm[L::JR + jc * S + s] = a;
    }
}

// World transforms level by level; also A_j^t = G_j^t - G_j^R J_j.
template <int S, class L = TileLayout<S>>
SB_HD void ph_chain_forward(const ModelView& M, float* sm, const Grp g) {
    for (int lev = 0; lev < M.num_levels; ++lev) {
        const int first = M.level_start[lev], cnt = M.level_start[lev + 1] - first;
        FOR_ITEMS_G(it, cnt * S, g) {
            const int s = it % S, j = M.level_order[first + it / S], p = M.parents[j];
            float G[12];
            const float Jx = sm[L::JR + (3 * j + 0) * S + s], Jy = sm[L::JR + (3 * j + 1) * S + s],
                        Jz = sm[L::JR + (3 * j + 2) * S + s];
            if (p < 0) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) G[r * 4 + c] = sm[L::RM + (j * 9 + r * 3 + c) * S + s];
                }
                G[3] = Jx; G[7] = Jy; G[11] = Jz;
            } else {
                float Rl[9], Gp[12];
#pragma unroll
                for (int e = 0; e < 9; ++e) Rl[e] = sm[L::RM + (j * 9 + e) * S + s];
#pragma unroll
                for (int e = 0; e < 12; ++e) Gp[e] = sm[L::GW + (p * 12 + e) * S + s];
                const float dx = Jx - sm[L::JR + (3 * p + 0) * S + s], dy = Jy - sm[L::JR + (3 * p + 1) * S + s],
                            dz = Jz - sm[L::JR + (3 * p + 2) * S + s];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        G[r * 4 + c] = Gp[r * 4 + 0] * Rl[c] + Gp[r * 4 + 1] * Rl[3 + c] + Gp[r * 4 + 2] * Rl[6 + c];
                    G[r * 4 + 3] = Gp[r * 4 + 0] * dx + Gp[r * 4 + 1] * dy + Gp[r * 4 + 2] * dz + Gp[r * 4 + 3];
                }
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) sm[L::GW + (j * 12 + e) * S + s] = G[e];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                sm[L::AT + (3 * j + r) * S + s] = G[r * 4 + 3] - (G[r * 4 + 0] * Jx + G[r * 4 + 1] * Jy + G[r * 4 + 2] * Jz);
        }
        grp_sync(g);
    }
}

SB_HD float4 ld_const4(const float4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// 4-column x (2 NP)-sample register tile, acc[c][s] += w[c] * v[s]  (NP = 4: 8 samples, NP = 3: 6 samples).  On the
// device the samples are packed in pairs and updated with Blackwell's packed fp32 FMA (fma.rn.f32x2 via __ffma2_rn):
// scalar FFMA issues at half rate on sm_100, the 2-wide form is what reaches the fp32 peak.  Results are bit-identical
// to scalar fmaf.
template <int NP>
struct TileAcc {
    static constexpr int NS = 2 * NP;
#if defined(__CUDA_ARCH__)
    float2 a[4][NP];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < NP; ++p) a[c][p] = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void fma(const float4& w, const float2 (&vp)[NP]) {
        const float2 wp[4] = {make_float2(w.x, w.x), make_float2(w.y, w.y), make_float2(w.z, w.z), make_float2(w.w, w.w)};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < NP; ++p) a[c][p] = __ffma2_rn(wp[c], vp[p], a[c][p]);
    }
    __device__ __forceinline__ float get(int c, int s) const { return (s & 1) ? a[c][s >> 1].y : a[c][s >> 1].x; }
#else
    float a[4][NS];
    void clear() {
        for (int c = 0; c < 4; ++c)
            for (int s = 0; s < NS; ++s) a[c][s] = 0.f;
    }
    void fma(const float4& w, const float2 (&vp)[NP]) {
        const float ww[4] = {w.x, w.y, w.z, w.w};
        for (int c = 0; c < 4; ++c)
            for (int p = 0; p < NP; ++p) {
                a[c][2 * p] = fmaf(ww[c], vp[p].x, a[c][2 * p]);
                a[c][2 * p + 1] = fmaf(ww[c], vp[p].y, a[c][2 * p + 1]);
            }
    }
    float get(int c, int s) const { return a[c][s]; }
#endif
    // dst[0 .. NS) += column c of the tile (dst 8-byte aligned): second half of a split reduction
    SB_HD void add_into(float* dst, int c) const {
        float2* o = reinterpret_cast<float2*>(dst);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float2 v = o[p];
            v.x += get(c, 2 * p);
            v.y += get(c, 2 * p + 1);
            o[p] = v;
        }
    }
    // column c of the tile -> dst[0 .. NS): 16-byte stores where the tile is 8 samples wide, 8-byte ones otherwise
    SB_HD void store(float* dst, int c, float sub = 0.f) const {
        if constexpr (NP == 4) {
            float4* o = reinterpret_cast<float4*>(dst);
            o[0] = make_float4(get(c, 0) - sub, get(c, 1) - sub, get(c, 2) - sub, get(c, 3) - sub);
            o[1] = make_float4(get(c, 4) - sub, get(c, 5) - sub, get(c, 6) - sub, get(c, 7) - sub);
        } else {
            float2* o = reinterpret_cast<float2*>(dst);
#pragma unroll
            for (int p = 0; p < NP; ++p) o[p] = make_float2(get(c, 2 * p) - sub, get(c, 2 * p + 1) - sub);
        }
    }
};

// Tile shapes of the GEMM fast paths: 8 samples per thread where S is a multiple of 8, 6 where S = 12 (the tile size that
// fills the second wave of batches such as 4096 = 148 x 16 + 144 x 12).
template <int S> constexpr bool kFastGemm = (S % 8 == 0) || (S == 12) || (S == 4);
template <int S> constexpr int kTileSamples = (S % 8 == 0) ? 8 : ((S == 12) ? 6 : 4);
// Tiles with one sample group (S = 4, 8) leave half of the 384 threads without a GEMM item, and a warp alone on its
// scheduler cannot hide the latency of the streamed constants: the device splits the reduction (K) dimension of the three
// GEMMs over the idle threads instead and adds the partial sums in a fixed order.
template <int S> constexpr bool kSplitK = kFastGemm<S> && (S / kTileSamples<S> == 1);

// One row of x for a tile: NP sample pairs starting at x (16-byte aligned when NP = 4, 8-byte aligned otherwise).
template <int NP>
SB_HD void load_x_row(const float* x, float2 (&v)[NP]) {
    if constexpr (NP == 4) {
        const float4 v0 = reinterpret_cast<const float4*>(x)[0], v1 = reinterpret_cast<const float4*>(x)[1];
        v[0] = make_float2(v0.x, v0.y); v[1] = make_float2(v0.z, v0.w); v[2] = make_float2(v1.x, v1.y); v[3] = make_float2(v1.z, v1.w);
    } else if constexpr (NP == 2) {
        const float4 v0 = reinterpret_cast<const float4*>(x)[0];
        v[0] = make_float2(v0.x, v0.y); v[1] = make_float2(v0.z, v0.w);
    } else {
#pragma unroll
        for (int p = 0; p < NP; ++p) v[p] = reinterpret_cast<const float2*>(x)[p];
    }
}

// acc[c][s] += sum_{k < K} w[k * ldw4][c] * x[k * ldx + s]  for the 4 columns of one float4 column quad and 2 NP samples:
// the streamed-constant GEMM core of the three GEMM phases.  The constant rows come from L2 in groups of 4, two groups
// ahead, through three register sets used in rotation; the x row of the next k-step is fetched from shared memory while
// the current one is multiplied.  Measured in isolation on B200 (tools/exp/gemm_probe.cu): 70 % of the fp32 FMA peak
// against 57 % for the plain two-deep prefetch loop.
template <int K, int LDW4, int LDX, int NP>
SB_HD void stream_gemm(TileAcc<NP>& acc, const float4* w, const float* x) {
    constexpr int U = 4, G = K / U;
    static_assert(K % U == 0 && G >= 2, "K must be a multiple of 4, at least 8");
    float4 ca[U], cb[U], cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { ca[u] = ld_const4(w + u * LDW4); cb[u] = ld_const4(w + (U + u) * LDW4); cc[u] = cb[u]; }
    float2 v[NP];
    load_x_row<NP>(x, v);
    // one group of U rows at (wg, xg), held in cur: fetch the group two ahead into nxt, multiply, keep x one row ahead.
    // All offsets from wg / xg are compile-time constants (immediate-offset loads, no per-load address arithmetic).
    auto step = [&](const float4 (&cur)[U], float4 (&nxt)[U], const float4* wg, const float* xg, bool prefetch, bool last) {
        if (prefetch) {
#pragma unroll
            for (int u = 0; u < U; ++u) nxt[u] = ld_const4(wg + (2 * U + u) * LDW4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int kn = (last && u == U - 1) ? u : u + 1;                     // next row (the very last one is re-read, unused)
            float2 n[NP];
            load_x_row<NP>(xg + kn * LDX, n);
            acc.fma(cur[u], v);
#pragma unroll
            for (int p = 0; p < NP; ++p) v[p] = n[p];
        }
    };
    constexpr int ROUNDS = (G - 2) / 3, TAIL = G - 3 * ROUNDS;          // TAIL in {2, 3, 4}
    constexpr int UNR = (ROUNDS % 3 == 0) ? 3 : ((ROUNDS <= 6) ? ROUNDS : ((ROUNDS % 2 == 0) ? 2 : 1));
    const float4* wg = w;
    const float* xg = x;
#pragma unroll UNR
    for (int r = 0; r < ROUNDS; ++r) {
        step(ca, cc, wg, xg, true, false);
        step(cb, ca, wg + U * LDW4, xg + U * LDX, true, false);
        step(cc, cb, wg + 2 * U * LDW4, xg + 2 * U * LDX, true, false);
        wg += 3 * U * LDW4;
        xg += 3 * U * LDX;
    }
    step(ca, cc, wg, xg, TAIL > 2, false);
    step(cb, ca, wg + U * LDW4, xg + U * LDX, TAIL > 3, TAIL == 2);
    if (TAIL > 2) step(cc, cb, wg + 2 * U * LDW4, xg + 2 * U * LDX, false, TAIL == 3);
    if (TAIL > 3) step(ca, cc, wg + 3 * U * LDW4, xg + 3 * U * LDX, false, true);
}

// Work item -> (column item p, sample group h) of a GEMM phase with H sample groups.  With two groups the two threads that
// share a column item sit in the same warp (lanes l and l + 16): their constant loads coalesce into one request, so every
// constant is fetched once per CTA instead of once per sample group (the streamed constants are the L2-bandwidth term).
template <int H>
SB_HD void gemm_item(int t, int& p, int& h) {
    static_assert(H == 1 || H == 2, "one or two sample groups");
    if (H == 2) { p = (t >> 5) * 16 + (t & 15); h = (t >> 4) & 1; }
    else { p = t; h = 0; }
}

#if defined(SMPLB200_EMU_TS) && !defined(__CUDA_ARCH__)
// TEST-ONLY numerics model of the tcgen05 3xTF32 path of the pair kernel (fit_pair.cuh), for the host emulation build:
// operands split by truncation into the 19 bits the tensor core reads (hi) and the truncated remainder (lo); every MMA adds
// the exact sum of its 8 products to the fp32 accumulator and TRUNCATES (what the tensor pipe was observed to do); hi.hi
// products go to a `main` accumulator, the lo.hi and hi.lo corrections to a second one; `parts` partial accumulator pairs over
// equal K ranges are added in fp32 at the end.
#include <cmath>
namespace tsemu {
inline float trunc19(float v) { union { float f; unsigned u; } c; c.f = v; c.u &= 0xFFFFE000u; return c.f; }
inline float trunc_f32(double v) { float f = (float)v; if (std::fabs((double)f) > std::fabs(v)) f = std::nextafterf(f, 0.f); return f; }
// a[k], b[k] for k < K (K a multiple of 8 after zero padding by the caller)
inline float dot3(const float* a, int astride, const float* b, int bstride, int K, int parts) {
    const int ksteps = (K + 7) / 8, per = (ksteps + parts - 1) / parts;
    float total = 0.f;
    for (int p = 0; p < parts; ++p) {
        float main = 0.f, corr = 0.f;
        for (int ks = p * per; ks < ksteps && ks < (p + 1) * per; ++ks) {
            double hh = 0, lh = 0, hl = 0;
            for (int k = ks * 8; k < ks * 8 + 8 && k < K; ++k) {
                const float av = a[(size_t)k * astride], bv = b[(size_t)k * bstride];
                const float ah = trunc19(av), al = trunc19(av - ah), bh = trunc19(bv), bl = trunc19(bv - bh);
                hh += (double)ah * bh; lh += (double)al * bh; hl += (double)ah * bl;
            }
            main = trunc_f32((double)main + hh);
            corr = trunc_f32((double)corr + lh);
            corr = trunc_f32((double)corr + hl);
        }
        total += main + corr;
    }
    return total;
}
}  // namespace tsemu
#endif

// QT[n][s] = sum_m Cf[m][n] * x[m][s]   (the folded joint GEMM: [S x 218] . [218 x 681])
// S % 8 == 0: one thread owns 4 adjacent columns x 8 samples (32 accumulators, 1 LDG.128 + 2 LDS.128 per
// 32 FMAs); the basis rows are streamed from L2 two groups of U rows ahead, x is broadcast from smem.
template <int S, class L = TileLayout<S>>
SB_HD_CALL void ph_fold_gemm_forward(const ModelView& M, float* sm) {
    static_assert(S % 4 == 0, "S must be a multiple of 4");
#if defined(SMPLB200_EMU_TS) && !defined(__CUDA_ARCH__)
    for (int n = 0; n < kQPad; ++n)
        for (int s = 0; s < S; ++s) sm[L::QT + n * L::LDQ + s] = tsemu::dot3(M.Cf + n, kQPad, sm + L::XT + s, S, kXPad, 1);
    return;
#endif
#if defined(__CUDA_ARCH__)
    if constexpr (kSplitK<S>) {
        // called by the 352 GEMM threads of the standard tile (the chain warp is busy elsewhere): thread = (column quad, K half)
        constexpr int NQ4 = kQPad / 4, KH = kXPad / 2;
        if (TILE_NT == kFitTileThreads) {
            const int t = TILE_TID, cq = t % NQ4, ks = t / NQ4;
            TileAcc<S / 2> acc;
            acc.clear();
            if (t < 2 * NQ4)
                stream_gemm<KH, NQ4, S>(acc, reinterpret_cast<const float4*>(M.Cf) + cq + (size_t)ks * KH * NQ4, sm + L::XT + ks * KH * S);
            if (t < NQ4) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc.store(sm + L::QT + (4 * cq + c) * L::LDQ, c);
            }
            if (t < 2 * NQ4) asm volatile("bar.sync 1, %0;" ::"n"(2 * NQ4) : "memory");       // the GEMM threads only
            if (t >= NQ4 && t < 2 * NQ4) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc.add_into(sm + L::QT + (4 * cq + c) * L::LDQ, c);
            }
            return;
        }
    }
#endif
    if constexpr (kFastGemm<S>) {
        constexpr int TS = kTileSamples<S>, NQ4 = kQPad / 4, H = S / TS;       // 176 column quads, H sample groups
        FOR_ITEMS(t, NQ4 * H) {
            int cq, h;
            gemm_item<H>(t, cq, h);
            TileAcc<TS / 2> acc;
            acc.clear();
            stream_gemm<kXPad, NQ4, S>(acc, reinterpret_cast<const float4*>(M.Cf) + cq, sm + L::XT + TS * h);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc.store(sm + L::QT + (4 * cq + c) * L::LDQ + TS * h, c);
        }
    } else {
        FOR_ITEMS(n, kQPad) {
            float a[S];
#pragma unroll
            for (int s = 0; s < S; ++s) a[s] = 0.f;
            for (int m = 0; m < kX; ++m) {
                const float c = M.Cf[m * kQPad + n];
#pragma unroll
                for (int s = 0; s < S; ++s) a[s] += c * sm[L::XT + m * S + s];
            }
#pragma unroll
            for (int s = 0; s < S; ++s) sm[L::QT + n * L::LDQ + s] = a[s];
        }
    }
}

// dx[m][s] = sum_n Cf[m][n] * dQ[n][s]   (transpose GEMM; result overwrites XT).
// Device fast path (S a multiple of 8 or S = 12, >= 384 threads): thread = (n-range r of 3, sample group h, column quad mq of
// 56), 4 x 8 accumulators; the three partial sums are combined in a fixed order through the (by then dead)
// QT region - deterministic, no atomics.
template <int S, class L = TileLayout<S>>
SB_HD_CALL void ph_fold_gemm_backward(const ModelView& M, float* sm) {
#if defined(SMPLB200_EMU_TS) && !defined(__CUDA_ARCH__)
    {
        float dx[kXPad * S];
        for (int m = 0; m < kXPad; ++m)
            for (int s = 0; s < S; ++s) dx[m * S + s] = tsemu::dot3(M.CfT + m, kXPad, sm + L::QT + s, L::LDQ, kQPad, SMPLB200_EMU_TS_BWD_PARTS);
        for (int i = 0; i < kXPad * S; ++i) sm[L::XT + i] = dx[i];
        return;
    }
#endif
#if defined(__CUDA_ARCH__)
    if constexpr (kFastGemm<S>) {
        // n ranges: 3 x 232 rows with two sample groups; with one group the idle threads take more, shorter ranges
        // (S = 8: 4 x 176, S = 4: 6 x 116; rows >= 681 of CfT are zero padding)
        constexpr int TS = kTileSamples<S>, MQ = kXPad / 4, H = S / TS;
        constexpr int NR = (H == 2) ? 3 : ((S == 8) ? 4 : 6), NPER = (H == 2) ? 232 : ((S == 8) ? 176 : 116);
        static_assert(NR * NPER <= kQPad && NR * NPER >= kQ && NPER % 4 == 0, "n ranges");
        static_assert((NR - 1) * kXPad * S <= kQPad * L::LDQ, "partial sums must fit in the QT region");
        if (TILE_NT >= MQ * NR * H) {
            int p, h;
            gemm_item<H>(TILE_TID, p, h);
            const int r = p / MQ, mq = p % MQ;
            const bool active = p < MQ * NR;                 // the chain warp (threads >= 352) maps to p >= 176: never active
            TileAcc<TS / 2> acc;
            acc.clear();
            if (active) {
                const int n_begin = r * NPER;
                stream_gemm<NPER, MQ, L::LDQ>(acc, reinterpret_cast<const float4*>(M.CfT) + (size_t)n_begin * MQ + mq,
                                              sm + L::QT + n_begin * L::LDQ + TS * h);
            }
            TILE_SYNC();                               // everyone is done reading dQ: QT becomes scratch
            if (active) {
                float* dst = (r == 0) ? (sm + L::XT) : (sm + L::QT + (r - 1) * kXPad * S);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc.store(dst + (4 * mq + c) * S + TS * h, c);
            }
            TILE_SYNC();
            FOR_ITEMS(i, kXPad * S) {
                float a = sm[L::XT + i];
#pragma unroll
                for (int r2 = 1; r2 < NR; ++r2) a += sm[L::QT + (r2 - 1) * kXPad * S + i];
                sm[L::XT + i] = a;
            }
            return;
        }
    }
#endif
    // generic path (host emulation, small S, or too few threads)
    FOR_ITEMS(m, kXPad) {
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = 0.f;
        const float* ct = M.CfT + m;
        for (int n = 0; n < kQ; ++n) {
            const float c = ct[n * kXPad];
#pragma unroll
            for (int s = 0; s < S; ++s) acc[s] += c * sm[L::QT + n * L::LDQ + s];
        }
#pragma unroll
        for (int s = 0; s < S; ++s) sm[L::XT + m * S + s] = acc[s];
    }
}

// The 20 "heavy" source joints - the 9 folded extra joints E_k = sum_j G_j^R Q_kj + A_j^t w_kj and the 11 skinned picked
// vertices P_p = sum_j Wp[p][j] (G_j^R v_p + A_j^t) - are sums over the 24 joints that all read the same G_j^R, A_j^t.
// A work item is (sample, quarter of the joints, block of <= 5 extras or <= 3 picks): the 12 shared transform entries of
// a joint are loaded once per item instead of once per (source, joint), and every source is computed once even when two
// outputs map to it.  The four partial positions per source go to a scratch area (the DG rows, dead until the
// reprojection phase clears them) and are added in a fixed order when the outputs are gathered.
constexpr int kHeavy = kExtra + kPicks;                 // 20 sources: picks 0..10, extras 11..19
constexpr int kJointParts = 4;                         // the 24 joints in 4 runs of 6
template <int S, class L = TileLayout<S>>
SB_HD void ph_output_joints(const ModelView& M, const SmallConsts& C, float* sm) {
    constexpr int JN = kJoints / kJointParts, EB = 5, PB = 3;
    constexpr int NEB = (kExtra + EB - 1) / EB, NPB = (kPicks + PB - 1) / PB;      // 2 extra blocks, 4 pick blocks
    static_assert(kHeavy * 3 * kJointParts <= 288, "partial positions must fit in the DG rows");
    float* scr = sm + L::DG;                              // [(heavy source * 3 + coordinate) * kJointParts + third][S]
    constexpr int PER_BLK = (kJointParts * S + 31) / 32 * 32;      // whole warps per block: extras and picks never share a warp
    FOR_ITEMS(it, (NEB + NPB) * PER_BLK) {
        const int blk = it / PER_BLK, idx = it % PER_BLK, s = idx % S, t = idx / S;
        if (t >= kJointParts) continue;
        if (blk < NEB) {
            const int k0 = blk * EB;
            float acc[EB][3];
#pragma unroll
            for (int i = 0; i < EB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
            for (int j = t * JN; j < (t + 1) * JN; ++j) {
                const float* G = sm + L::GW + (j * 12) * S + s;
                const float g0 = G[0 * S], g1 = G[1 * S], g2 = G[2 * S], g4 = G[4 * S], g5 = G[5 * S], g6 = G[6 * S],
                            g8 = G[8 * S], g9 = G[9 * S], g10 = G[10 * S];
                const float t0 = sm[L::AT + (3 * j + 0) * S + s], t1 = sm[L::AT + (3 * j + 1) * S + s], t2 = sm[L::AT + (3 * j + 2) * S + s];
#pragma unroll
                for (int i = 0; i < EB; ++i) {
                    const int k = k0 + i;
                    if (k < kExtra) {
                        const int qb = (k * kJoints + j) * 3;
                        const float qx = sm[L::q(qb + 0, s)], qy = sm[L::q(qb + 1, s)], qz = sm[L::q(qb + 2, s)];
                        const float w = C.wkj[k * kJoints + j];
                        acc[i][0] += g0 * qx + g1 * qy + g2 * qz + t0 * w;
                        acc[i][1] += g4 * qx + g5 * qy + g6 * qz + t1 * w;
                        acc[i][2] += g8 * qx + g9 * qy + g10 * qz + t2 * w;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < EB; ++i)
                if (k0 + i < kExtra) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) scr[(((kPicks + k0 + i) * 3 + c) * kJointParts + t) * S + s] = acc[i][c];
                }
        } else {
            const int p0 = (blk - NEB) * PB;
            float acc[PB][3], v[PB][3];
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
                const int p = (p0 + i < kPicks) ? p0 + i : kPicks - 1;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[i][c] = sm[L::q(kQPickBase + 3 * p + c, s)];
            }
            for (int j = t * JN; j < (t + 1) * JN; ++j) {
                const float* G = sm + L::GW + (j * 12) * S + s;
                const float g0 = G[0 * S], g1 = G[1 * S], g2 = G[2 * S], g4 = G[4 * S], g5 = G[5 * S], g6 = G[6 * S],
                            g8 = G[8 * S], g9 = G[9 * S], g10 = G[10 * S];
                const float t0 = sm[L::AT + (3 * j + 0) * S + s], t1 = sm[L::AT + (3 * j + 1) * S + s], t2 = sm[L::AT + (3 * j + 2) * S + s];
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    const int p = (p0 + i < kPicks) ? p0 + i : kPicks - 1;
                    const float w = C.Wp[p * kJoints + j];
                    acc[i][0] += w * (g0 * v[i][0] + g1 * v[i][1] + g2 * v[i][2] + t0);
                    acc[i][1] += w * (g4 * v[i][0] + g5 * v[i][1] + g6 * v[i][2] + t1);
                    acc[i][2] += w * (g8 * v[i][0] + g9 * v[i][1] + g10 * v[i][2] + t2);
                }
            }
#pragma unroll
            for (int i = 0; i < PB; ++i)
                if (p0 + i < kPicks) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) scr[(((p0 + i) * 3 + c) * kJointParts + t) * S + s] = acc[i][c];
                }
        }
    }
    TILE_SYNC();
    // gather the 49 outputs: a chain joint is the translation of its world transform, a heavy source the sum of its partials
    FOR_ITEMS(it, kOut * S) {
        const int s = it % S, o = it / S, src = M.joint_map[o];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v;
            if (src < kJoints) {
                v = sm[L::GW + (src * 12 + 4 * c + 3) * S + s];
            } else {
                const int h = (src < kJoints + kSelVerts) ? src - kJoints : kPicks + (src - (kJoints + kSelVerts));   // picks < kPicks only
                const float* q = scr + ((h * 3 + c) * kJointParts) * S + s;
                v = ((q[0] + q[S]) + q[2 * S]) + q[3 * S];
            }
            sm[L::OUTJ + (3 * o + c) * S + s] = v;
        }
    }
}

// Reprojection term of body_fitting_loss per output joint; with_grad also overwrites OUTJ with
// dL/djoint.  LOSSJ[o] = conf^2 * (gmof(u - kx) + gmof(v - ky)).
template <int S, class L = TileLayout<S>>
SB_HD void ph_reprojection(float* sm, float focal, float sigma2, bool with_grad) {
    FOR_ITEMS(it, kOut * S) {
        const int s = it % S, o = it / S;
        const float Px = sm[L::OUTJ + (3 * o + 0) * S + s] + sm[L::CAM + 0 * S + s];
        const float Py = sm[L::OUTJ + (3 * o + 1) * S + s] + sm[L::CAM + 1 * S + s];
        const float Pz = sm[L::OUTJ + (3 * o + 2) * S + s] + sm[L::CAM + 2 * S + s];
        const float px = Px / Pz, py = Py / Pz;
        const float du = (focal * px + sm[L::CEN + 0 * S + s]) - sm[L::KP + (3 * o + 0) * S + s];
        const float dv = (focal * py + sm[L::CEN + 1 * S + s]) - sm[L::KP + (3 * o + 1) * S + s];
        const float conf = sm[L::KP + (3 * o + 2) * S + s];
        const float c2 = conf * conf;
        sm[L::LOSSJ + o * S + s] = c2 * (gmof_val(du, sigma2) + gmof_val(dv, sigma2));
        if (with_grad) {
            const float gu = c2 * gmof_grad(du, sigma2) * focal, gv = c2 * gmof_grad(dv, sigma2) * focal;
            sm[L::OUTJ + (3 * o + 0) * S + s] = gu / Pz;
            sm[L::OUTJ + (3 * o + 1) * S + s] = gv / Pz;
            sm[L::OUTJ + (3 * o + 2) * S + s] = -(gu * px + gv * py) / Pz;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// priors (smplify/prior.py:181-196 merged max-mixture; losses.py:19-24 angle prior; shape prior)
// ------------------------------------------------------------------------------------------------
// Pd[(g,i)][s] = sum_j Psym_g[i][j] * bp[j][s] - (Psym_g mean_g)[i]   for all 8 components: an
// [S x 72] . [72 x 576] GEMM streamed like the folded one (two adjacent i per thread).
template <int S, class L = TileLayout<S>>
SB_HD_CALL void ph_prior_quadratic(const ModelView& M, const SmallConsts& C, float* sm) {
#if defined(SMPLB200_EMU_TS) && !defined(__CUDA_ARCH__)
    for (int g = 0; g < kGauss; ++g)
        for (int i = 0; i < kPriorPad; ++i)
            for (int s = 0; s < S; ++s)
                sm[L::QT + (g * kPriorPad + i) * S + s] =
                    tsemu::dot3(M.gmm_prec + (size_t)g * kPriorPad * kPriorPad + i, kPriorPad, sm + L::POSE + 3 * S + s, S, kPriorPad, 1) -
                    C.pmean[g * kPriorPad + i];
    return;
#endif
#if defined(__CUDA_ARCH__)
    if constexpr (kSplitK<S>) {
        constexpr int IQ = kPriorPad / 4, KH = kPriorPad / 2, NI = kGauss * IQ;        // 144 (component, column quad) items x 2 K halves
        if (TILE_NT >= 2 * NI) {
            const int t = TILE_TID, gi = t % NI, ks = t / NI, g = gi / IQ, iq = gi % IQ;
            TileAcc<S / 2> acc;
            acc.clear();
            if (t < 2 * NI)
                stream_gemm<KH, IQ, S>(acc, reinterpret_cast<const float4*>(M.gmm_prec + (size_t)g * kPriorPad * kPriorPad) + iq + (size_t)ks * KH * IQ,
                                       sm + L::POSE + (3 + ks * KH) * S);
            if (t < NI) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc.store(sm + L::QT + (g * kPriorPad + 4 * iq + c) * S, c, C.pmean[g * kPriorPad + 4 * iq + c]);
            }
            TILE_SYNC();
            if (t >= NI && t < 2 * NI) {
#pragma unroll
                for (int c = 0; c < 4; ++c) acc.add_into(sm + L::QT + (g * kPriorPad + 4 * iq + c) * S, c);
            }
            return;
        }
    }
#endif
    if constexpr (kFastGemm<S>) {
        // 4 adjacent i x 8 (or 6) samples per thread through the streamed-constant GEMM core (rows j >= 69 of Psym are
        // zero; the pose rows they meet - the first betas - are finite)
        constexpr int TS = kTileSamples<S>, IQ = kPriorPad / 4, H = S / TS;        // 18 column quads per component
        FOR_ITEMS(t, kGauss * IQ * H) {
            int gi, h;
            gemm_item<H>(t, gi, h);
            const int g = gi / IQ, iq = gi % IQ;
            TileAcc<TS / 2> acc;
            acc.clear();
            stream_gemm<kPriorPad, IQ, S>(acc, reinterpret_cast<const float4*>(M.gmm_prec + (size_t)g * kPriorPad * kPriorPad) + iq,
                                          sm + L::POSE + 3 * S + TS * h);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                acc.store(sm + L::QT + (g * kPriorPad + 4 * iq + c) * S + TS * h, c, C.pmean[g * kPriorPad + 4 * iq + c]);
        }
    } else {
        constexpr int U = 8, IP = kPriorPad / 2;        // 36 column pairs per component
        static_assert(kPriorPad % U == 0, "prior padding");
        FOR_ITEMS(t, kGauss * IP) {
            const int g = t / IP, ip = t % IP;
            float a0[S], a1[S];
    #pragma unroll
            for (int s = 0; s < S; ++s) { a0[s] = 0.f; a1[s] = 0.f; }
            const float2* P = reinterpret_cast<const float2*>(M.gmm_prec + (size_t)g * kPriorPad * kPriorPad) + ip;
            float2 cur[U], nxt[U];
    #pragma unroll
            for (int u = 0; u < U; ++u) { cur[u] = ld_const2(P + u * IP); nxt[u] = cur[u]; }
    #pragma unroll 1
            for (int j0 = 0; j0 < kPriorPad; j0 += U) {
                if (j0 + U < kPriorPad) {
    #pragma unroll
                    for (int u = 0; u < U; ++u) nxt[u] = ld_const2(P + (j0 + U + u) * IP);
                }
    #pragma unroll
                for (int u = 0; u < U; ++u) {
                    // rows j >= 69 of Psym are zero; the pose rows they meet (first betas) are finite
                    const float4* br = reinterpret_cast<const float4*>(sm + L::POSE + (3 + j0 + u) * S);
    #pragma unroll
                    for (int q = 0; q < S / 4; ++q) {
                        const float4 v = br[q];
                        a0[4 * q + 0] += cur[u].x * v.x; a0[4 * q + 1] += cur[u].x * v.y;
                        a0[4 * q + 2] += cur[u].x * v.z; a0[4 * q + 3] += cur[u].x * v.w;
                        a1[4 * q + 0] += cur[u].y * v.x; a1[4 * q + 1] += cur[u].y * v.y;
                        a1[4 * q + 2] += cur[u].y * v.z; a1[4 * q + 3] += cur[u].y * v.w;
                    }
                }
    #pragma unroll
                for (int u = 0; u < U; ++u) cur[u] = nxt[u];
            }
            const float pm0 = C.pmean[g * kPriorPad + 2 * ip], pm1 = C.pmean[g * kPriorPad + 2 * ip + 1];
            float* o = sm + L::QT + (g * kPriorPad + 2 * ip) * S;
    #pragma unroll
            for (int s = 0; s < S; ++s) { o[s] = a0[s] - pm0; o[S + s] = a1[s] - pm1; }
        }
    }
}

// QSPLIT > 1 (small tiles, where 8 S items leave most threads idle): every quadratic form is summed in QSPLIT runs of
// consecutive terms by QSPLIT threads and the partial sums are added in run order (through the GPR rows, written last).
template <int S, class L = TileLayout<S>, int QSPLIT = 1>
SB_HD void ph_prior_select(const ModelView& M, const SmallConsts& C, float* sm, float prior_w2, float angle_w2, float shape_w2,
                           const Grp grp = grp_tile()) {
    if constexpr (QSPLIT > 1) {
        constexpr int RUN = (kPriorDim + QSPLIT - 1) / QSPLIT;
        static_assert(QSPLIT * kGauss <= kPriorDim, "partial sums must fit in the GPR rows");
        FOR_ITEMS_G(it, QSPLIT * kGauss * S, grp) {
            const int s = it % S, g = (it / S) % kGauss, part = it / (kGauss * S);
            const int i1 = (part + 1) * RUN < kPriorDim ? (part + 1) * RUN : kPriorDim;
            float q = 0.f;
            for (int i = part * RUN; i < i1; ++i)
                q += sm[L::pd(g, i, s)] * (sm[L::POSE + (3 + i) * S + s] - C.mu[g * kPriorPad + i]);
            sm[L::GPR + (part * kGauss + g) * S + s] = q;
        }
        grp_sync(grp);
        FOR_ITEMS_G(it, kGauss * S, grp) {
            const int s = it % S, g = it / S;
            float q = sm[L::GPR + g * S + s];
#pragma unroll
            for (int part = 1; part < QSPLIT; ++part) q += sm[L::GPR + (part * kGauss + g) * S + s];
            sm[L::MISC + g * S + s] = 0.5f * q - C.lognll[g];
        }
    } else {
        FOR_ITEMS_G(it, kGauss * S, grp) {
            const int s = it % S, g = it / S;
            float q = 0.f;
            for (int i = 0; i < kPriorDim; ++i)
                q += sm[L::pd(g, i, s)] * (sm[L::POSE + (3 + i) * S + s] - C.mu[g * kPriorPad + i]);
            sm[L::MISC + g * S + s] = 0.5f * q - C.lognll[g];
        }
    }
    grp_sync(grp);
    FOR_ITEMS_G(s, S, grp) {
        int best = 0;
        float bv = sm[L::MISC + s];
        for (int g = 1; g < kGauss; ++g) {
            const float v = sm[L::MISC + g * S + s];
            if (v < bv) { bv = v; best = g; }
        }
        sm[L::MISC + 8 * S + s] = (float)best;
        sm[L::LOSSJ + 49 * S + s] = prior_w2 * bv;
        float ang = 0.f;
        for (int a = 0; a < 4; ++a) {
            const float e = expf(sm[L::POSE + (3 + M.angle_ids[a]) * S + s] * M.angle_signs[a]);
            ang += e * e;
        }
        sm[L::LOSSJ + 50 * S + s] = angle_w2 * ang;
        float sh = 0.f;
        for (int l = 0; l < kBetas; ++l) sh += sm[L::BETA + l * S + s] * sm[L::BETA + l * S + s];
        sm[L::LOSSJ + 51 * S + s] = shape_w2 * sh;
    }
    grp_sync(grp);
    // gradient of the selected component: prior_w2 * Psym (bp - mean); plus the angle prior
    FOR_ITEMS_G(it, kPriorDim * S, grp) {
        const int s = it % S, i = it / S;
        const int g = (int)sm[L::MISC + 8 * S + s];
        float gr = prior_w2 * sm[L::pd(g, i, s)];
        for (int k = 0; k < 4; ++k)
            if (M.angle_ids[k] == i) {
                const float e = expf(sm[L::POSE + (3 + i) * S + s] * M.angle_signs[k]);
                gr += angle_w2 * 2.0f * M.angle_signs[k] * e * e;
            }
        sm[L::GPR + i * S + s] = gr;
    }
}

// ------------------------------------------------------------------------------------------------
// backward phases
// ------------------------------------------------------------------------------------------------
template <int S, class L = TileLayout<S>>
SB_HD void source_grad(const ModelView& M, const float* sm, int src, int s, float* d) {
    d[0] = d[1] = d[2] = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int o = M.inv_map[src][t];
        if (o >= 0) {
            d[0] += sm[L::OUTJ + (3 * o + 0) * S + s];
            d[1] += sm[L::OUTJ + (3 * o + 1) * S + s];
            d[2] += sm[L::OUTJ + (3 * o + 2) * S + s];
        }
    }
}

// Gradients arriving at the 20 heavy source joints (kHeavy: picks 0..10, extras 11..19), gathered once per sample into the
// XT rows - x is dead between the forward and the backward folded GEMM - instead of once per (joint, source).
template <int S, class L = TileLayout<S>>
SB_HD void ph_source_grads(const ModelView& M, float* sm) {
    static_assert(kHeavy * 3 <= kXPad, "source gradients must fit in the XT rows");
    FOR_ITEMS(it, kHeavy * S) {
        const int s = it % S, h = it / S;
        const int src = (h < kPicks) ? kJoints + h : kJoints + kSelVerts + (h - kPicks);
        float d[3];
        source_grad<S, L>(M, sm, src, s, d);
#pragma unroll
        for (int c = 0; c < 3; ++c) sm[L::XT + (3 * h + c) * S + s] = d[c];
    }
}

// dL/dA_j from the extra joints and the picked vertices, converted to dL/dG_j and dL/dJ_j.  DG holds additional dL/dA
// rows on entry (zero in the fit, the vertex path's dA in SMPL backward).  Starts with ph_source_grads + a tile barrier.
template <int S, class L = TileLayout<S>>
SB_HD void ph_joint_backward(const ModelView& M, const SmallConsts& C, float* sm) {
    ph_source_grads<S, L>(M, sm);
    TILE_SYNC();
    const float* SG = sm + L::XT;
    FOR_ITEMS(it, kJoints * S) {
        const int s = it % S, j = it / S;
        float dAR[9], dAt[3];
#pragma unroll
        for (int e = 0; e < 9; ++e) dAR[e] = sm[L::DG + (j * 12 + (e / 3) * 4 + e % 3) * S + s];
#pragma unroll
        for (int r = 0; r < 3; ++r) dAt[r] = sm[L::DG + (j * 12 + r * 4 + 3) * S + s];
        float G[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) G[r * 3 + c] = sm[L::GW + (j * 12 + r * 4 + c) * S + s];
        for (int k = 0; k < kExtra; ++k) {
            const float dE[3] = {SG[(3 * (kPicks + k) + 0) * S + s], SG[(3 * (kPicks + k) + 1) * S + s], SG[(3 * (kPicks + k) + 2) * S + s]};
            const int qb = (k * kJoints + j) * 3;
            const float qx = sm[L::q(qb + 0, s)], qy = sm[L::q(qb + 1, s)], qz = sm[L::q(qb + 2, s)];
            const float w = C.wkj[k * kJoints + j];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                dAR[r * 3 + 0] += dE[r] * qx; dAR[r * 3 + 1] += dE[r] * qy; dAR[r * 3 + 2] += dE[r] * qz;
                dAt[r] += w * dE[r];
            }
            L::store_dq(sm, qb + 0, s, G[0] * dE[0] + G[3] * dE[1] + G[6] * dE[2]);
            L::store_dq(sm, qb + 1, s, G[1] * dE[0] + G[4] * dE[1] + G[7] * dE[2]);
            L::store_dq(sm, qb + 2, s, G[2] * dE[0] + G[5] * dE[1] + G[8] * dE[2]);
        }
        for (int p = 0; p < kPicks; ++p) {
            const float dV[3] = {SG[(3 * p + 0) * S + s], SG[(3 * p + 1) * S + s], SG[(3 * p + 2) * S + s]};
            const float w = C.Wp[p * kJoints + j];
            const float vx = sm[L::q(kQPickBase + 3 * p + 0, s)], vy = sm[L::q(kQPickBase + 3 * p + 1, s)],
                        vz = sm[L::q(kQPickBase + 3 * p + 2, s)];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float wd = w * dV[r];
                dAR[r * 3 + 0] += wd * vx; dAR[r * 3 + 1] += wd * vy; dAR[r * 3 + 2] += wd * vz;
                dAt[r] += wd;
            }
        }
        // A^R = G^R ; A^t = G^t - G^R J
        const float Jx = sm[L::JR + (3 * j + 0) * S + s], Jy = sm[L::JR + (3 * j + 1) * S + s], Jz = sm[L::JR + (3 * j + 2) * S + s];
        float dJt[3];
        source_grad<S, L>(M, sm, j, s, dJt);           // the chain joint itself is G_j^t
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            sm[L::DG + (j * 12 + r * 4 + 0) * S + s] = dAR[r * 3 + 0] - dAt[r] * Jx;
            sm[L::DG + (j * 12 + r * 4 + 1) * S + s] = dAR[r * 3 + 1] - dAt[r] * Jy;
            sm[L::DG + (j * 12 + r * 4 + 2) * S + s] = dAR[r * 3 + 2] - dAt[r] * Jz;
            sm[L::DG + (j * 12 + r * 4 + 3) * S + s] = dAt[r] + dJt[r];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            sm[L::DJ + (3 * j + c) * S + s] = -(G[0 + c] * dAt[0] + G[3 + c] * dAt[1] + G[6 + c] * dAt[2]);
    }
}

// dL/d(v_posed of picked vertex) = (sum_j Wp[p][j] G_j^R)^T dL/dvert ; written over the pick rows of QT.  Reads the source
// gradients ph_joint_backward left in XT; runs after it (a tile barrier in between: it overwrites the pick rows that phase reads).
template <int S, class L = TileLayout<S>>
SB_HD void ph_pick_backward(const ModelView& M, const SmallConsts& C, float* sm) {
    const float* SG = sm + L::XT;
    FOR_ITEMS(it, kPicks * S) {
        const int s = it % S, p = it / S;
        const float dV[3] = {SG[(3 * p + 0) * S + s], SG[(3 * p + 1) * S + s], SG[(3 * p + 2) * S + s]};
        float T[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) T[e] = 0.f;
        for (int j = 0; j < kJoints; ++j) {
            const float w = C.Wp[p * kJoints + j];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) T[r * 3 + c] += w * sm[L::GW + (j * 12 + r * 4 + c) * S + s];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            L::store_dq(sm, kQPickBase + 3 * p + c, s, T[0 + c] * dV[0] + T[3 + c] * dV[1] + T[6 + c] * dV[2]);
    }
}

// Reverse sweep of the kinematic chain.  Parents gather from their children level by level
// (deterministic, no atomics); then every joint derives dL/dR_j (stored over RM) and the
// rest-joint gradient DJ.
template <int S, class L = TileLayout<S>>
SB_HD void ph_chain_backward(const ModelView& M, float* sm, const Grp g) {
    for (int lev = M.num_levels - 2; lev >= 0; --lev) {
        const int first = M.level_start[lev], cnt = M.level_start[lev + 1] - first;
        FOR_ITEMS_G(it, cnt * S, g) {
            const int s = it % S, p = M.level_order[first + it / S];
            const int c0 = M.child_start[p], c1 = M.child_start[p + 1];
            if (c0 == c1) continue;
            float dGp[12], Gp[9], dJp[3];
#pragma unroll
            for (int e = 0; e < 12; ++e) dGp[e] = sm[L::DG + (p * 12 + e) * S + s];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Gp[r * 3 + c] = sm[L::GW + (p * 12 + r * 4 + c) * S + s];
#pragma unroll
            for (int c = 0; c < 3; ++c) dJp[c] = sm[L::DJ + (3 * p + c) * S + s];
            const float Jpx = sm[L::JR + (3 * p + 0) * S + s], Jpy = sm[L::JR + (3 * p + 1) * S + s], Jpz = sm[L::JR + (3 * p + 2) * S + s];
            for (int ci = c0; ci < c1; ++ci) {
                const int i = M.child_list[ci];
                float dGi[12], Ri[9];
#pragma unroll
                for (int e = 0; e < 12; ++e) dGi[e] = sm[L::DG + (i * 12 + e) * S + s];
#pragma unroll
                for (int e = 0; e < 9; ++e) Ri[e] = sm[L::RM + (i * 9 + e) * S + s];
                const float rel[3] = {sm[L::JR + (3 * i + 0) * S + s] - Jpx, sm[L::JR + (3 * i + 1) * S + s] - Jpy,
                                      sm[L::JR + (3 * i + 2) * S + s] - Jpz};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        dGp[r * 4 + c] += dGi[r * 4 + 0] * Ri[c * 3 + 0] + dGi[r * 4 + 1] * Ri[c * 3 + 1] + dGi[r * 4 + 2] * Ri[c * 3 + 2]
                                          + dGi[r * 4 + 3] * rel[c];
                    dGp[r * 4 + 3] += dGi[r * 4 + 3];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)      // d(J_i - J_p) = G_p^R^T dG_i^t ; parent gets the minus sign
                    dJp[c] -= Gp[0 + c] * dGi[3] + Gp[3 + c] * dGi[7] + Gp[6 + c] * dGi[11];
            }
#pragma unroll
            for (int e = 0; e < 12; ++e) sm[L::DG + (p * 12 + e) * S + s] = dGp[e];
#pragma unroll
            for (int c = 0; c < 3; ++c) sm[L::DJ + (3 * p + c) * S + s] = dJp[c];
        }
        grp_sync(g);
    }
    FOR_ITEMS_G(it, kJoints * S, g) {
        const int s = it % S, j = it / S, p = M.parents[j];
        float dGi[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) dGi[e] = sm[L::DG + (j * 12 + e) * S + s];
        if (p < 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) sm[L::RM + (j * 9 + r * 3 + c) * S + s] = dGi[r * 4 + c];
#pragma unroll
            for (int c = 0; c < 3; ++c) sm[L::DJ + (3 * j + c) * S + s] += dGi[c * 4 + 3];
        } else {
            float Gp[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Gp[r * 3 + c] = sm[L::GW + (p * 12 + r * 4 + c) * S + s];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)   // dR_j = G_p^R^T dG_j^R
                    sm[L::RM + (j * 9 + r * 3 + c) * S + s] = Gp[0 + r] * dGi[0 + c] + Gp[3 + r] * dGi[4 + c] + Gp[6 + r] * dGi[8 + c];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                sm[L::DJ + (3 * j + c) * S + s] += Gp[0 + c] * dGi[3] + Gp[3 + c] * dGi[7] + Gp[6 + c] * dGi[11];
        }
    }
}

// dL/dR_j of body joints also receives the pose-feature gradient dx[11 + 9(j-1) + e].
template <int S, class L = TileLayout<S>>
SB_HD void rotation_grad(const float* sm, int j, int s, float* g) {
#pragma unroll
    for (int e = 0; e < 9; ++e) {
        g[e] = sm[L::RM + (j * 9 + e) * S + s];
        if (j > 0) g[e] += sm[L::XT + (11 + (j - 1) * 9 + e) * S + s];
    }
}

// dL/dbeta_l = dx[1+l] + sum_{j,c} JS[j][c][l] dJ[j][c]  (the prior term is added by the caller)
template <int S, class L = TileLayout<S>>
SB_HD float beta_grad(const SmallConsts& C, const float* sm, int l, int s) {
    float a = sm[L::XT + (1 + l) * S + s];
    for (int jc = 0; jc < 72; ++jc) a += C.JS[jc * kBetas + l] * sm[L::DJ + jc * S + s];
    return a;
}

template <int S, class L = TileLayout<S>>
SB_HD void zero_rows(float* sm, int off, int rows) {
    FOR_ITEMS(it, rows * S) sm[off + it] = 0.f;
}

}  // namespace smplb200
