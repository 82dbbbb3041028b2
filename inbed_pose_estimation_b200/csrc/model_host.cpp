// Model-constant construction.  The linear maps the reference applies per call in fp32
//   J        = J_regressor . (v_template + shapedirs.beta)                      (smplx lbs, SURVEY §8a a6,a7)
//   extra_k  = J_regressor_extra . verts,  verts = (W.A)[v_posed;1]             (models/smpl.py:24, a11)
//   picked_p = verts[vertex_id_p]                                               (VertexJointSelector, a12)
// are composed once here in float64:
//   J0 = Jr.v_template, JS = Jr.shapedirs
//   Cf[(k,j,c)][m] = sum_v Jx[k,v] W[v,j] basis[m][3v+c],  wkj[k][j] = sum_v Jx[k,v] W[v,j]
//   Cf[(p,c)][m]   = basis[m][3 vid_p + c],                Wp[p][j]  = W[vid_p][j]
// so that every joint the loss needs is an exact algebraic function of x = [1, beta, pose_feature]
// and the 24 skinning transforms, without touching the 6890 vertices inside the fit loop.
#include "model_host.h"
#include "lbs_tc.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <thread>

namespace smplb200 {

static void parallel_for(int n, const std::function<void(int)>& fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
    nt = std::min(nt, n);
    if (nt <= 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=, &fn]() {
            for (int i = t; i < n; i += nt) fn(i);
        });
    for (auto& t : th) t.join();
}

std::string build_host_model(const smplb200_model_desc& d, HostModel& H) {
    if (!d.v_template || !d.shapedirs || !d.posedirs || !d.J_regressor || !d.weights || !d.J_regressor_extra ||
        !d.parents || !d.extra_vertex_ids || !d.joint_map)
        return "model description has a NULL required array";
    ModelView& V = H.view;
    memset(&V, 0, sizeof(V));

    // ---- integer tables -----------------------------------------------------------------------
    if (d.parents[0] >= 0) return "parents[0] must be negative (root)";
    int depth[kJoints];
    depth[0] = 0;
    V.parents[0] = -1;
    int max_depth = 0;
    for (int j = 1; j < kJoints; ++j) {
        const int p = d.parents[j];
        if (p < 0 || p >= j) return "parents must satisfy 0 <= parents[j] < j for j >= 1";
        V.parents[j] = (int8_t)p;
        depth[j] = depth[p] + 1;
        max_depth = std::max(max_depth, depth[j]);
    }
    V.num_levels = max_depth + 1;
    int pos = 0;
    for (int lev = 0; lev <= max_depth; ++lev) {
        V.level_start[lev] = (uint8_t)pos;
        for (int j = 0; j < kJoints; ++j)
            if (depth[j] == lev) V.level_order[pos++] = (uint8_t)j;
    }
    V.level_start[max_depth + 1] = (uint8_t)pos;
    pos = 0;
    for (int p = 0; p < kJoints; ++p) {
        V.child_start[p] = (uint8_t)pos;
        for (int j = 1; j < kJoints; ++j)
            if (d.parents[j] == p) V.child_list[pos++] = (uint8_t)j;
    }
    V.child_start[kJoints] = (uint8_t)pos;
    for (int s = 0; s < kSrc; ++s) V.inv_map[s][0] = V.inv_map[s][1] = -1;
    for (int o = 0; o < kOut; ++o) {
        const int s = d.joint_map[o];
        if (s < 0 || s >= kSrc) return "joint_map entry out of range [0,54)";
        if (s >= kJoints + kPicks && s < kJoints + kSelVerts)
            return "joint_map references a selected-vertex joint beyond the 11 the folded model carries (35..44)";
        V.joint_map[o] = (uint8_t)s;
        if (V.inv_map[s][0] < 0) V.inv_map[s][0] = (int8_t)o;
        else if (V.inv_map[s][1] < 0) V.inv_map[s][1] = (int8_t)o;
        else return "a source joint feeds more than two outputs";
    }
    for (int p = 0; p < kSelVerts; ++p) {
        if (d.extra_vertex_ids[p] < 0 || d.extra_vertex_ids[p] >= kVerts) return "extra_vertex_ids out of range";
        V.pick_vid[p] = d.extra_vertex_ids[p];
    }
    V.num_ign = 0;
    if (d.ign_joints) {
        if (d.num_ign_joints < 0 || d.num_ign_joints > 8) return "num_ign_joints must be in [0,8]";
        V.num_ign = d.num_ign_joints;
        for (int i = 0; i < V.num_ign; ++i) {
            if (d.ign_joints[i] < 0 || d.ign_joints[i] >= kOut) return "ign_joints out of range";
            V.ign_joints[i] = (uint8_t)d.ign_joints[i];
        }
    }
    for (int q = 0; q < 4; ++q) {
        const int op = d.cam_op_joints ? d.cam_op_joints[q] : 0, gt = d.cam_gt_joints ? d.cam_gt_joints[q] : 0;
        if (op < 0 || op >= kOut || gt < 0 || gt >= kOut) return "camera joints out of range";
        V.cam_op[q] = (uint8_t)op;
        V.cam_gt[q] = (uint8_t)gt;
        const int id = d.angle_prior_ids ? d.angle_prior_ids[q] : 0;
        if (id < 0 || id >= kPriorDim) return "angle_prior_ids out of range";
        V.angle_ids[q] = id;
        V.angle_signs[q] = d.angle_prior_signs ? d.angle_prior_signs[q] : 0.f;
    }

    // ---- blend basis (row m of x -> 20670 coordinates) and its transpose -------------------------
    H.basis.assign((size_t)kXPad * kColsPad, 0.f);
    H.basisT.assign((size_t)kColsPad * kXPad, 0.f);
    for (int n = 0; n < kCols; ++n) {
        H.basis[n] = d.v_template[n];
        for (int l = 0; l < kBetas; ++l) H.basis[(size_t)(1 + l) * kColsPad + n] = d.shapedirs[(size_t)n * kBetas + l];
        for (int f = 0; f < kPoseFeat; ++f) H.basis[(size_t)(11 + f) * kColsPad + n] = d.posedirs[(size_t)n * kPoseFeat + f];
    }
    for (int m = 0; m < kX; ++m)
        for (int n = 0; n < kCols; ++n) H.basisT[(size_t)n * kXPad + m] = H.basis[(size_t)m * kColsPad + n];
    H.weights.assign((size_t)(kVerts + 64) * kJoints, 0.f);
    memcpy(H.weights.data(), d.weights, sizeof(float) * (size_t)kVerts * kJoints);

    // ---- tf32 hi/lo split of the tensor-core operands (3xTF32) ------------------------------------------
    auto split = [](const std::vector<float>& src, std::vector<float>& hi, std::vector<float>& lo) {
        hi.assign(src.size(), 0.f);
        lo.assign(src.size(), 0.f);
        for (size_t i = 0; i < src.size(); ++i) {
            hi[i] = tf32_round(src[i]);
            lo[i] = src[i] - hi[i];
        }
    };
    split(H.basisT, H.basisT_hi, H.basisT_lo);
    split(H.basis, H.basis_hi, H.basis_lo);
    {
        std::vector<float> w32((size_t)kTcVertRowsPad * 32, 0.f), wT((size_t)32 * kTcVertRowsPad, 0.f);
        for (int v = 0; v < kVerts; ++v)
            for (int j = 0; j < kJoints; ++j) {
                const float w = d.weights[(size_t)v * kJoints + j];
                w32[(size_t)v * 32 + j] = w;
                wT[(size_t)j * kTcVertRowsPad + v] = w;
            }
        split(w32, H.w_hi, H.w_lo);
        split(wT, H.wT_hi, H.wT_lo);
    }

    // ---- rest joints: J0 + JS.beta ------------------------------------------------------------------
    H.J0.assign(72, 0.f);
    H.JS.assign(72 * kBetas, 0.f);
    for (int j = 0; j < kJoints; ++j)
        for (int c = 0; c < 3; ++c) {
            double a0 = 0.0, as[kBetas] = {0};
            for (int v = 0; v < kVerts; ++v) {
                const double r = d.J_regressor[(size_t)j * kVerts + v];
                a0 += r * d.v_template[3 * v + c];
                for (int l = 0; l < kBetas; ++l) as[l] += r * d.shapedirs[(size_t)(3 * v + c) * kBetas + l];
            }
            H.J0[3 * j + c] = (float)a0;
            for (int l = 0; l < kBetas; ++l) H.JS[(3 * j + c) * kBetas + l] = (float)as[l];
        }

    // ---- folded extra joints and picked vertices --------------------------------------------------------
    H.Cf.assign((size_t)kXPad * kQPad, 0.f);
    H.CfT.assign((size_t)kQPad * kXPad, 0.f);
    H.wkj.assign(kExtra * kJoints, 0.f);
    H.Wp.assign(kPicks * kJoints, 0.f);
    std::vector<double> wsum(kExtra * kJoints, 0.0);
    const float* basis = H.basis.data();
    parallel_for(kExtra * kJoints, [&](int kj) {
        const int k = kj / kJoints, j = kj % kJoints;
        std::vector<double> coef(kVerts);
        double ws = 0.0;
        for (int v = 0; v < kVerts; ++v) {
            coef[v] = (double)d.J_regressor_extra[(size_t)k * kVerts + v] * (double)d.weights[(size_t)v * kJoints + j];
            ws += coef[v];
        }
        wsum[kj] = ws;
        for (int m = 0; m < kX; ++m) {
            const float* row = basis + (size_t)m * kColsPad;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int v = 0; v < kVerts; ++v) {
                a0 += coef[v] * row[3 * v]; a1 += coef[v] * row[3 * v + 1]; a2 += coef[v] * row[3 * v + 2];
            }
            const double a[3] = {a0, a1, a2};
            for (int c = 0; c < 3; ++c) {
                const int n = kj * 3 + c;
                H.Cf[(size_t)m * kQPad + n] = (float)a[c];
                H.CfT[(size_t)n * kXPad + m] = (float)a[c];
            }
        }
    });
    for (int kj = 0; kj < kExtra * kJoints; ++kj) H.wkj[kj] = (float)wsum[kj];
    for (int p = 0; p < kPicks; ++p) {
        const int vid = V.pick_vid[p];
        for (int j = 0; j < kJoints; ++j) H.Wp[p * kJoints + j] = d.weights[(size_t)vid * kJoints + j];
        for (int c = 0; c < 3; ++c) {
            const int n = kQPickBase + 3 * p + c;
            for (int m = 0; m < kX; ++m) {
                const float b = basis[(size_t)m * kColsPad + 3 * vid + c];
                H.Cf[(size_t)m * kQPad + n] = b;
                H.CfT[(size_t)n * kXPad + m] = b;
            }
        }
    }
    for (int s = 0; s < kSrc; ++s) V.sigma_src[s] = 1.f;
    for (int p = 0; p < kPicks; ++p) {
        double a = 0.0;
        for (int j = 0; j < kJoints; ++j) a += H.Wp[p * kJoints + j];
        V.sigma_src[kJoints + p] = (float)a;
    }
    for (int k = 0; k < kExtra; ++k) {
        double a = 0.0;
        for (int j = 0; j < kJoints; ++j) a += wsum[k * kJoints + j];
        V.sigma_src[kJoints + kSelVerts + k] = (float)a;
    }

    // ---- max-mixture prior ------------------------------------------------------------------------------
    H.has_prior = d.gmm_means && d.gmm_precisions && d.gmm_nll_weights;
    H.gmm_means.assign(kGauss * kPriorPad, 0.f);
    H.gmm_prec.assign((size_t)kGauss * kPriorPad * kPriorPad, 0.f);
    H.gmm_pmean.assign(kGauss * kPriorPad, 0.f);
    H.gmm_lognll.assign(kGauss, 0.f);
    if (H.has_prior) {
        for (int g = 0; g < kGauss; ++g) {
            for (int i = 0; i < kPriorDim; ++i) H.gmm_means[g * kPriorPad + i] = d.gmm_means[g * kPriorDim + i];
            const float* P = d.gmm_precisions + (size_t)g * kPriorDim * kPriorDim;
            float* Ps = H.gmm_prec.data() + (size_t)g * kPriorPad * kPriorPad;
            for (int i = 0; i < kPriorDim; ++i)
                for (int j = 0; j < kPriorDim; ++j)   // d^T P d == d^T Psym d ; gradient 0.5 (P + P^T) d == Psym d
                    Ps[j * kPriorPad + i] = (float)(0.5 * ((double)P[i * kPriorDim + j] + (double)P[j * kPriorDim + i]));
            for (int i = 0; i < kPriorDim; ++i) {
                double a = 0.0;
                for (int j = 0; j < kPriorDim; ++j) a += (double)Ps[j * kPriorPad + i] * (double)d.gmm_means[g * kPriorDim + j];
                H.gmm_pmean[g * kPriorPad + i] = (float)a;
            }
            if (!(d.gmm_nll_weights[g] > 0.f)) return "gmm_nll_weights must be positive (representable in fp32)";
            H.gmm_lognll[g] = (float)log((double)d.gmm_nll_weights[g]);
        }
    }
    // ---- packed A operands of the pair kernel's tensor-core GEMMs -----------------------------------------------------
    // element (row R of the tile space, k) of a GEMM with KC chunks sits at float index
    //   ((((t * 2 + h) * KC + k / 32) * 8 + (k % 32) / 4) * 128 + r) * 4 + k % 4,    R = 256 t + 128 h + r
    auto pack = [](std::vector<float>& out, int tiles, int kchunks, int rows, int K, const std::function<float(int, int)>& at) {
        out.assign((size_t)tiles * 2 * kchunks * kPgChunk * 128, 0.f);
        for (int R = 0; R < rows; ++R) {
            const int t = R / 256, h = (R % 256) / 128, r = R % 128;
            for (int k = 0; k < K; ++k) {
                const size_t idx = ((((size_t)(t * 2 + h) * kchunks + k / kPgChunk) * 8 + (k % kPgChunk) / 4) * 128 + r) * 4 + k % 4;
                out[idx] = at(R, k);
            }
        }
    };
    pack(H.pg_fwd, kPgFwdTiles, kPgFwdChunks, kQPad, kXPad, [&](int n, int m) { return H.Cf[(size_t)m * kQPad + n]; });
    pack(H.pg_bwd, kPgBwdTiles, kPgBwdChunks, kXPad, kQPad, [&](int m, int n) { return H.Cf[(size_t)m * kQPad + n]; });
    pack(H.pg_prior, kPgPriorTiles, kPgPriorChunks, kGauss * kPriorPad, kPriorPad, [&](int R, int j) {
        const int g = R / kPriorPad, i = R % kPriorPad;
        return H.gmm_prec[(size_t)g * kPriorPad * kPriorPad + (size_t)j * kPriorPad + i];
    });
    return std::string();
}

}  // namespace smplb200
