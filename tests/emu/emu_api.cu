// TEST-ONLY host emulation of the tile code (fit_tile / pose_forward_tile / pose_backward_tile).
// The phases in csrc/fit_tile.cuh are __host__ __device__; compiled here for the host they run
// with one "thread" per tile (TILE_TID = 0, TILE_NT = 1, no barrier), which is a valid
// serialisation because all cross-phase state lives in the tile's scratch array.  This lets the
// math (forward, hand-derived backward, priors, Adam) be checked against the oracle on a box
// without a GPU.  It is NOT part of the product: nothing under inbed_pose_estimation_b200/
// loads this library, and the product library has no host execution path.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../inbed_pose_estimation_b200/csrc/fit_driver.cuh"
#include "../../inbed_pose_estimation_b200/csrc/model_host.h"

using namespace smplb200;

struct EmuModel {
    HostModel H;
};
static std::string g_err;

extern "C" const char* emu_last_error() { return g_err.c_str(); }

extern "C" EmuModel* emu_model_create(const smplb200_model_desc* d) {
    EmuModel* m = new EmuModel();
    g_err = build_host_model(*d, m->H);
    if (!g_err.empty()) { delete m; return nullptr; }
    ModelView& V = m->H.view;
    HostModel& H = m->H;
    V.basis = H.basis.data(); V.basisT = H.basisT.data(); V.weights = H.weights.data();      // unused by the tile code
    V.Cf = H.Cf.data(); V.CfT = H.CfT.data(); V.wkj = H.wkj.data(); V.Wp = H.Wp.data();
    V.J0 = H.J0.data(); V.JS = H.JS.data();
    V.gmm_means = H.gmm_means.data(); V.gmm_prec = H.gmm_prec.data(); V.gmm_pmean = H.gmm_pmean.data();
    V.gmm_lognll = H.gmm_lognll.data();
    return m;
}
extern "C" void emu_model_destroy(EmuModel* m) { delete m; }

constexpr int S = 8;

// hi/lo tf32 operands of the tcgen05 vertex kernels, as the tile code writes them; recombined (exactly) into the plain
// A [B][24][12] and x [B][224] the tests look at
struct Operands {
    std::vector<float> x_hi, x_lo, ae_hi, ae_lo;
    TcOperands tc;
    explicit Operands(int batch)
        : x_hi((size_t)batch * kXPad), x_lo((size_t)batch * kXPad), ae_hi((size_t)batch * kAeRow), ae_lo((size_t)batch * kAeRow) {
        tc.x_hi = x_hi.data(); tc.x_lo = x_lo.data(); tc.ae_hi = ae_hi.data(); tc.ae_lo = ae_lo.data();
    }
    void unpack(int batch, float* A, float* x) const {
        if (x)
            for (size_t i = 0; i < (size_t)batch * kXPad; ++i) x[i] = x_hi[i] + x_lo[i];
        if (A)
            for (int b = 0; b < batch; ++b)
                for (int j = 0; j < kJoints; ++j)
                    for (int e = 0; e < 12; ++e) {
                        const size_t o = (size_t)b * kAeRow + e * 32 + j;
                        A[(size_t)b * 288 + j * 12 + e] = ae_hi[o] + ae_lo[o];
                    }
    }
};

static std::vector<float> tile_scratch() { return std::vector<float>(TileLayout<S>::SMEM_FLOATS + 64, 0.f); }

extern "C" int emu_fit(EmuModel* m, int batch, int num_iters, double step_size, float focal, int loss_only,
                       const float* pose, const float* betas, const float* cam, const float* center, float* kp,
                       float* joints, float* opose, float* obetas, float* ocam, float* reproj, float* trace,
                       float* ws_A, float* ws_x) {
    FitParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch; P.num_iters = loss_only ? 0 : num_iters; P.zero_conf_first = loss_only; P.focal = focal;
    P.init_pose = pose; P.init_betas = betas; P.init_cam = cam; P.center = center; P.keypoints = kp;
    P.out_joints = joints; P.out_pose = opose; P.out_betas = obetas; P.out_cam = ocam; P.out_reproj = reproj;
    Operands ops(batch);
    P.tc = ops.tc;
    P.loss_trace = trace;
    P.lr = step_size; P.beta1 = 0.9; P.beta2 = 0.999;
    P.adam_c.lerp_w = (float)(1.0 - 0.9); P.adam_c.beta2 = (float)0.999; P.adam_c.w2 = (float)(1.0 - 0.999); P.adam_c.eps = 1e-8f;
    std::vector<float> sm = tile_scratch();
    for (int t = 0; t < (batch + S - 1) / S; ++t) fit_tile<S>(m->H.view, P, t * S, sm.data());
    ops.unpack(batch, ws_A, ws_x);
    return 0;
}

extern "C" int emu_pose(EmuModel* m, int batch, int rotmat_mode, int backward, const float* pose, const float* betas,
                        float* joints, float* ws_A, float* ws_x, const float* d_joints, const float* dA_part,
                        const float* dx_part, int nsplit, float* d_pose, float* d_betas) {
    PoseParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch; P.rotmat_mode = rotmat_mode; P.pose = pose; P.betas = betas; P.joints = joints;
    Operands ops(batch);
    P.tc = ops.tc;
    // the tests hand dL/dA in as [B][24][12]; the dA kernel's layout is entry-major [B][12][24]
    std::vector<float> dA_e;
    if (dA_part) {
        dA_e.resize((size_t)batch * 288);
        for (int b = 0; b < batch; ++b)
            for (int j = 0; j < kJoints; ++j)
                for (int e = 0; e < 12; ++e) dA_e[((size_t)b * 12 + e) * kJoints + j] = dA_part[(size_t)b * 288 + j * 12 + e];
    }
    P.d_joints = d_joints; P.dA_part = dA_part ? dA_e.data() : nullptr; P.dx_part = dx_part; P.nsplit_a = nsplit; P.nsplit_x = nsplit;
    P.d_pose = d_pose; P.d_betas = d_betas;
    std::vector<float> sm = tile_scratch();
    for (int t = 0; t < (batch + S - 1) / S; ++t) {
        if (backward) pose_backward_tile<S>(m->H.view, P, t * S, sm.data());
        else pose_forward_tile<S>(m->H.view, P, t * S, sm.data());
    }
    if (!backward) ops.unpack(batch, ws_A, ws_x);
    return 0;
}

extern "C" int emu_prior(EmuModel* m, int batch, const float* pose, const float* betas, float* terms, float* components, int* argmin,
                         float* grad_body_pose, float* grad_betas) {
    PriorParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch; P.pose = pose; P.betas = betas; P.terms = terms; P.components = components; P.argmin = argmin;
    P.grad_body_pose = grad_body_pose; P.grad_betas = grad_betas;
    std::vector<float> sm = tile_scratch();
    for (int t = 0; t < (batch + S - 1) / S; ++t) prior_tile<S>(m->H.view, P, t * S, sm.data());
    return 0;
}

// raw access for constant-folding checks
extern "C" const float* emu_model_array(EmuModel* m, const char* name, int* n) {
    HostModel& H = m->H;
    std::vector<float>* v = nullptr;
    std::string s(name);
    if (s == "Cf") v = &H.Cf; else if (s == "J0") v = &H.J0; else if (s == "JS") v = &H.JS;
    else if (s == "wkj") v = &H.wkj; else if (s == "Wp") v = &H.Wp; else if (s == "gmm_prec") v = &H.gmm_prec;
    else if (s == "gmm_lognll") v = &H.gmm_lognll; else if (s == "basis") v = &H.basis;
    if (!v) return nullptr;
    *n = (int)v->size();
    return v->data();
}
