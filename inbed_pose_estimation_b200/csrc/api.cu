// extern "C" boundary of libsmplify_b200.so (see include/smplify_b200.h).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/smplify_b200.h"
#include "fit_driver.cuh"
#include "launch.h"
#include "lbs_tc.h"
#include "model_host.h"
#include <stdlib.h>

using namespace smplb200;

struct smplb200_model {
    int device = 0;
    ModelView view;
    bool has_prior = false;
    TcConstMaps tc_maps;           // TMA maps of the tcgen05 vertex kernels' constant operands
    std::vector<void*> allocations;
    // grow-only device scratch + pinned staging of the host-buffer entry point
    std::mutex mu;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaStream_t host_stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device -> host copies of the fit results overlap the vertex kernels
    cudaEvent_t fit_done = nullptr;
};

static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};      // process-wide: autograd runs backward calls on its own threads

static int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}
#define CUDA_OK(expr)                                                                  \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) return fail("%s: %s", #expr, cudaGetErrorString(_e));   \
    } while (0)

// Makes `device` current for the scope and restores the caller's current device on every exit path: the entry points
// that own a device (create / destroy / the host-buffer fit) must not leave a different device current behind them.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

extern "C" int smplb200_version(void) { return 200; }
extern "C" void smplb200_fit_tile_plan(int batch, int sms, int* n16, int* small, int* n_small) {
    int a = 0, b = 0, c = 0;
    if (batch > 0 && sms > 0) plan_fit_tiles(batch, sms, &a, &b, &c);
    if (n16) *n16 = a;
    if (small) *small = b;
    if (n_small) *n_small = c;
}
extern "C" int smplb200_fit_pair_plan(int batch, int sms, int* n16, int* n12) {
    int a = 0, b = 0;
    const int pairs = fit_uses_pairs(batch, 1);
    if (pairs && sms > 1) plan_fit_pairs(batch, sms, &a, &b);
    if (n16) *n16 = a;
    if (n12) *n12 = b;
    return pairs;
}
extern "C" int smplb200_fit_split_plan(int batch, int sms) {
    if (batch <= 0) return 0;
    if (sms <= 0) return plan_fit_split_device(batch);            // the current device's cluster occupancy: what the launch uses
    return sms > 1 ? plan_fit_split(batch, sms) : 0;
}
extern "C" const char* smplb200_last_error(void) { return g_error.c_str(); }
#if defined(SMPLB200_PHASE_CLOCKS)
// profiling builds only (tools/phase_clocks.py): read / reset the stage-2 phase cycle counters
namespace smplb200 { cudaError_t debug_phase_clocks(unsigned long long* out32, int reset); }
extern "C" int smplb200_debug_phase_clocks(unsigned long long* out32, int reset) {
    return smplb200::debug_phase_clocks(out32, reset) == cudaSuccess ? 0 : 1;
}
#endif

#if defined(PG_TRACE)
namespace smplb200 { cudaError_t debug_pg_trace(long long* out512); }
extern "C" int smplb200_debug_pg_trace(long long* out512) { return smplb200::debug_pg_trace(out512) == cudaSuccess ? 0 : 1; }
#endif

extern "C" long long smplb200_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

template <typename T>
static int upload(smplb200_model* m, const std::vector<T>& h, const T** out) {
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, h.size() * sizeof(T)));
    m->allocations.push_back(p);
    CUDA_OK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(p);
    return 0;
}

extern "C" int smplb200_model_create(const smplb200_model_desc* desc, int device, smplb200_model** out) {
    if (!desc || !out) return fail("smplb200_model_create: NULL argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("smplb200_model_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("smplb200_model_create: device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail("smplb200_model_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    HostModel H;
    const std::string err = build_host_model(*desc, H);
    if (!err.empty()) return fail("smplb200_model_create: %s", err.c_str());
    DeviceGuard guard(device);
    CUDA_OK(guard.err);
    smplb200_model* m = new smplb200_model();
    m->device = device;
    m->view = H.view;
    m->has_prior = H.has_prior;
    int rc = 0;
    rc |= upload(m, H.Cf, &m->view.Cf);
    rc |= upload(m, H.CfT, &m->view.CfT);
    rc |= upload(m, H.wkj, &m->view.wkj);
    rc |= upload(m, H.Wp, &m->view.Wp);
    rc |= upload(m, H.J0, &m->view.J0);
    rc |= upload(m, H.JS, &m->view.JS);
    rc |= upload(m, H.gmm_means, &m->view.gmm_means);
    rc |= upload(m, H.gmm_prec, &m->view.gmm_prec);
    rc |= upload(m, H.gmm_pmean, &m->view.gmm_pmean);
    rc |= upload(m, H.gmm_lognll, &m->view.gmm_lognll);
    rc |= upload(m, H.pg_prior, &m->view.pg_prior);
    rc |= upload(m, H.pg_fwd, &m->view.pg_fwd);
    rc |= upload(m, H.pg_bwd, &m->view.pg_bwd);
    const float *bt_hi = nullptr, *bt_lo = nullptr, *bm_hi = nullptr, *bm_lo = nullptr, *w_hi = nullptr, *w_lo = nullptr,
                *wT_hi = nullptr, *wT_lo = nullptr;
    rc |= upload(m, H.basisT_hi, &bt_hi);
    rc |= upload(m, H.basisT_lo, &bt_lo);
    rc |= upload(m, H.basis_hi, &bm_hi);
    rc |= upload(m, H.basis_lo, &bm_lo);
    rc |= upload(m, H.w_hi, &w_hi);
    rc |= upload(m, H.w_lo, &w_lo);
    rc |= upload(m, H.wT_hi, &wT_hi);
    rc |= upload(m, H.wT_lo, &wT_lo);
    if (!rc) {
        memset(&m->tc_maps, 0, sizeof(m->tc_maps));
        if (!tc_make_constant_maps(&m->tc_maps, bt_hi, bt_lo, bm_hi, bm_lo, w_hi, w_lo, wT_hi, wT_lo))
            rc = fail("smplb200_model_create: cuTensorMapEncodeTiled failed (TMA descriptors of the tcgen05 kernels)");
    }
    if (rc) {
        smplb200_model_destroy(m);
        return 1;
    }
    *out = m;
    return 0;
}

extern "C" void smplb200_model_destroy(smplb200_model* m) {
    if (!m) return;
    DeviceGuard guard(m->device);
    for (void* p : m->allocations) cudaFree(p);
    if (m->scratch) cudaFree(m->scratch);
    if (m->host_stream) cudaStreamDestroy(m->host_stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->fit_done) cudaEventDestroy(m->fit_done);
    delete m;
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// Device workspace.  The pose kernels write the per-sample operands of the tcgen05 vertex kernels (x hi/lo [B][224],
// skinning transforms hi/lo [B][12][32]); `big0` holds v_posed [B][20736] in forward calls without a caller-supplied
// saved buffer and dvp (fp32) in backward calls, `big2` the padded copy of dverts (backward only), then
// the per-split partial sums.
struct Work {
    TcOperands tc;
    float *big0, *big2, *dA, *dx;
};
static size_t big_bytes(int batch) { return align256((size_t)batch * kVpPitch * 4); }
static size_t work_bytes(int batch, bool with_backward) {
    const size_t x = align256((size_t)batch * kXPad * 4), ae = align256((size_t)batch * kAeRow * 4);
    size_t n = 2 * x + 2 * ae + big_bytes(batch) + 256;
    if (with_backward)
        n += big_bytes(batch) + (size_t)kMaxSplitA * align256((size_t)batch * 288 * 4) + (size_t)kMaxSplitX * x;
    return n;
}
static Work carve(void* ws, int batch) {
    char* w = reinterpret_cast<char*>(align256(reinterpret_cast<size_t>(ws)));
    const size_t x = align256((size_t)batch * kXPad * 4), ae = align256((size_t)batch * kAeRow * 4);
    Work k;
    k.tc.x_hi = reinterpret_cast<float*>(w); w += x;
    k.tc.x_lo = reinterpret_cast<float*>(w); w += x;
    k.tc.ae_hi = reinterpret_cast<float*>(w); w += ae;
    k.tc.ae_lo = reinterpret_cast<float*>(w); w += ae;
    k.big0 = reinterpret_cast<float*>(w); w += big_bytes(batch);
    k.big2 = reinterpret_cast<float*>(w); w += big_bytes(batch);          // backward workspaces only (from here on)
    k.dA = reinterpret_cast<float*>(w); w += (size_t)kMaxSplitA * align256((size_t)batch * 288 * 4);
    k.dx = reinterpret_cast<float*>(w);
    return k;
}

extern "C" size_t smplb200_fit_workspace_bytes(int batch) { return batch < 0 ? 0 : work_bytes(batch, false); }
extern "C" size_t smplb200_smpl_workspace_bytes(int batch) { return batch < 0 ? 0 : work_bytes(batch, true); }

// blend-shape GEMM + skinning of all 6890 vertices from the operands the pose / fit kernel left in the workspace
static int run_vertices(const smplb200_model* m, const Work& wk, float* vposed, float* vertices, int batch, cudaStream_t st) {
    CUDA_OK(launch_blend_gemm(m->tc_maps, wk.tc.x_hi, wk.tc.x_lo, vposed, batch, st));
    CUDA_OK(launch_skin_forward(m->tc_maps, wk.tc.ae_hi, wk.tc.ae_lo, vposed, vertices, batch, st));
    g_launches += 2;
    return 0;
}

static AdamConsts adam_consts(double beta1, double beta2) {
    AdamConsts c;
    c.lerp_w = (float)(1.0 - beta1);
    c.beta2 = (float)beta2;
    c.w2 = (float)(1.0 - beta2);
    c.eps = 1e-8f;
    return c;
}

static int run_fit(const smplb200_model* m, int batch, int num_iters, double step_size, float focal, int loss_only,
                   const float* pose, const float* betas, const float* cam, const float* center, float* kp,
                   float* vertices, float* joints, float* opose, float* obetas, float* ocam, float* reproj, float* trace,
                   void* ws, size_t ws_bytes, cudaStream_t st, cudaEvent_t after_fit = nullptr, float* packed = nullptr) {
    if (!m) return fail("NULL model");
    if (batch < 0) return fail("negative batch");
    if (batch == 0) return 0;
    if (!loss_only && !m->has_prior) return fail("smplify_fit: the model was created without a GMM prior");
    if (num_iters < 0) return fail("num_iters must not be negative");
    if (!pose || !betas || !cam || !center || !kp || !reproj) return fail("smplify: NULL required buffer");
    if (ws_bytes < smplb200_fit_workspace_bytes(batch) || !ws) return fail("smplify: workspace too small");
    const Work wk = carve(ws, batch);
    FitParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch;
    P.num_iters = loss_only ? 0 : num_iters;
    P.zero_conf_first = loss_only;
    P.focal = focal;
    P.init_pose = pose; P.init_betas = betas; P.init_cam = cam; P.center = center; P.keypoints = kp;
    P.out_joints = joints; P.out_pose = opose; P.out_betas = obetas; P.out_cam = ocam; P.out_reproj = reproj;
    P.out_packed = packed;
    if (vertices) P.tc = wk.tc;
    P.loss_trace = trace;
    P.lr = step_size; P.beta1 = 0.9; P.beta2 = 0.999;        // lr stays a double: torch divides the Python float by bc1
    P.adam_c = adam_consts(P.beta1, P.beta2);
    CUDA_OK(launch_fit(m->view, P, st));
    ++g_launches;
    if (after_fit) CUDA_OK(cudaEventRecord(after_fit, st));          // parameters / joints / losses are final here
    if (vertices && run_vertices(m, wk, wk.big0, vertices, batch, st)) return 1;
    return 0;
}

extern "C" int smplb200_smplify_fit(const smplb200_model* model, int batch, int num_iters, double step_size, float focal_length,
                                    const float* init_pose, const float* init_betas, const float* init_cam_t,
                                    const float* camera_center, float* keypoints_2d, float* vertices, float* joints,
                                    float* pose, float* betas, float* camera_translation, float* reprojection_loss,
                                    float* loss_trace, float* packed_results, void* workspace, size_t workspace_bytes, void* stream) {
    if (!joints || !pose || !betas || !camera_translation) return fail("smplify_fit: NULL output buffer");
    return run_fit(model, batch, num_iters, step_size, focal_length, 0, init_pose, init_betas, init_cam_t, camera_center,
                   keypoints_2d, vertices, joints, pose, betas, camera_translation, reprojection_loss, loss_trace, workspace,
                   workspace_bytes, static_cast<cudaStream_t>(stream), nullptr, packed_results);
}

extern "C" int smplb200_smplify_fitting_loss(const smplb200_model* model, int batch, float focal_length, const float* pose,
                                             const float* betas, const float* cam_t, const float* camera_center,
                                             float* keypoints_2d, float* reprojection_loss, void* workspace,
                                             size_t workspace_bytes, void* stream) {
    return run_fit(model, batch, 0, 0.0, focal_length, 1, pose, betas, cam_t, camera_center, keypoints_2d, nullptr, nullptr,
                   nullptr, nullptr, nullptr, reprojection_loss, nullptr, workspace, workspace_bytes,
                   static_cast<cudaStream_t>(stream));
}

extern "C" int smplb200_prior_terms(const smplb200_model* m, int batch, const float* pose, const float* betas, float* terms,
                                    float* components, int32_t* argmin, float* grad_body_pose, float* grad_betas, void* stream) {
    if (!m) return fail("NULL model");
    if (!m->has_prior) return fail("prior_terms: the model was created without a GMM prior");
    if (batch < 0) return fail("negative batch");
    if (batch == 0) return 0;
    if (!pose || !betas || !terms) return fail("prior_terms: NULL required buffer");
    PriorParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch; P.pose = pose; P.betas = betas; P.terms = terms; P.components = components;
    P.argmin = reinterpret_cast<int*>(argmin); P.grad_body_pose = grad_body_pose; P.grad_betas = grad_betas;
    CUDA_OK(launch_prior_terms(m->view, P, static_cast<cudaStream_t>(stream)));
    ++g_launches;
    return 0;
}

extern "C" int smplb200_smpl_forward(const smplb200_model* m, int batch, int rotmat_mode, const float* pose, const float* betas,
                                     float* vertices, float* joints, float* saved_vposed, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (!m) return fail("NULL model");
    if (batch < 0) return fail("negative batch");
    if (batch == 0) return 0;
    if (!pose || !betas) return fail("smpl_forward: NULL input");
    if (!workspace || workspace_bytes < smplb200_smpl_workspace_bytes(batch)) return fail("smpl_forward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Work wk = carve(workspace, batch);
    PoseParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch; P.rotmat_mode = rotmat_mode; P.pose = pose; P.betas = betas; P.joints = joints;
    if (vertices) P.tc = wk.tc;
    CUDA_OK(launch_pose_forward(m->view, P, st));
    ++g_launches;
    if (vertices && run_vertices(m, wk, saved_vposed ? saved_vposed : wk.big0, vertices, batch, st)) return 1;
    return 0;
}

extern "C" int smplb200_smpl_backward(const smplb200_model* m, int batch, int rotmat_mode, const float* pose, const float* betas,
                                      const float* saved_vposed, const float* grad_vertices, const float* grad_joints,
                                      float* grad_pose, float* grad_betas, void* workspace, size_t workspace_bytes, void* stream) {
    if (!m) return fail("NULL model");
    if (batch < 0) return fail("negative batch");
    if (batch == 0) return 0;
    if (!pose || !betas || !grad_pose || !grad_betas) return fail("smpl_backward: NULL buffer");
    if (grad_vertices && !saved_vposed) return fail("smpl_backward: grad_vertices needs saved_vposed from the forward call");
    if (!workspace || workspace_bytes < smplb200_smpl_workspace_bytes(batch)) return fail("smpl_backward: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Work wk = carve(workspace, batch);
    PoseParams P;
    memset(&P, 0, sizeof(P));
    P.batch = batch; P.rotmat_mode = rotmat_mode; P.pose = pose; P.betas = betas;
    if (grad_vertices) {
        // The skinning transforms of the forward pass are recomputed (cheap) so that the backward call is self-contained.
        PoseParams F = P;
        F.tc = wk.tc;
        F.tc.x_hi = nullptr; F.tc.x_lo = nullptr;
        CUDA_OK(launch_pose_forward(m->view, F, st));
        const int nsa = tc_dA_splits(batch), nsx = tc_dx_splits(batch);
        CUDA_OK(launch_skin_backward(m->tc_maps, wk.tc.ae_hi, wk.tc.ae_lo, grad_vertices, wk.big0, wk.big2, batch, st));
        CUDA_OK(launch_dx_gemm(m->tc_maps, wk.big0, wk.dx, batch, nsx, st));
        CUDA_OK(launch_dA(m->tc_maps, wk.big2, saved_vposed, wk.dA, batch, nsa, st));
        g_launches += 4;
        P.dA_part = wk.dA; P.dx_part = wk.dx; P.nsplit_a = nsa; P.nsplit_x = nsx;
    }
    P.d_joints = grad_joints; P.d_pose = grad_pose; P.d_betas = grad_betas;
    CUDA_OK(launch_pose_backward(m->view, P, st));
    ++g_launches;
    return 0;
}

extern "C" int smplb200_batch_rodrigues(int n, const float* theta, float* rotmat, void* stream) {
    if (n < 0 || (n > 0 && (!theta || !rotmat))) return fail("batch_rodrigues: bad arguments");
    CUDA_OK(launch_quat_rodrigues_fwd(theta, rotmat, n, static_cast<cudaStream_t>(stream)));
    if (n) ++g_launches;
    return 0;
}
extern "C" int smplb200_batch_rodrigues_backward(int n, const float* theta, const float* grad_rotmat, float* grad_theta, void* stream) {
    if (n < 0 || (n > 0 && (!theta || !grad_rotmat || !grad_theta))) return fail("batch_rodrigues_backward: bad arguments");
    CUDA_OK(launch_quat_rodrigues_bwd(theta, grad_rotmat, grad_theta, n, static_cast<cudaStream_t>(stream)));
    if (n) ++g_launches;
    return 0;
}

extern "C" int smplb200_perspective_projection(int batch, int num_points, const float* points, const float* rotation,
                                               const float* translation, const float* focal_length, int focal_per_batch,
                                               const float* camera_center, int out_3d, float* projected, void* stream) {
    if (batch < 0 || num_points < 0) return fail("perspective_projection: negative size");
    if ((size_t)batch * num_points == 0) return 0;
    if (!points || !rotation || !translation || !focal_length || !camera_center || !projected)
        return fail("perspective_projection: NULL buffer");
    CUDA_OK(launch_projection_fwd(points, rotation, translation, focal_length, focal_per_batch, camera_center, projected, out_3d != 0,
                                  batch, num_points, static_cast<cudaStream_t>(stream)));
    ++g_launches;
    return 0;
}
extern "C" int smplb200_perspective_projection_backward(int batch, int num_points, const float* points, const float* rotation,
                                                        const float* translation, const float* focal_length, int focal_per_batch,
                                                        int out_3d, const float* grad_projected, float* grad_points,
                                                        float* grad_rotation, float* grad_translation, void* stream) {
    if (batch < 0 || num_points < 0) return fail("perspective_projection_backward: negative size");
    if (batch == 0) return 0;
    if (!points || !rotation || !translation || !focal_length || !grad_projected || !grad_points || !grad_rotation || !grad_translation)
        return fail("perspective_projection_backward: NULL buffer");
    CUDA_OK(launch_projection_bwd(points, rotation, translation, focal_length, focal_per_batch, grad_projected, out_3d != 0, grad_points,
                                  grad_rotation, grad_translation, batch, num_points, static_cast<cudaStream_t>(stream)));
    ++g_launches;
    return 0;
}

extern "C" int smplb200_weak_perspective_projection(int batch, int num_points, const float* joints, const float* pred_camera,
                                                    float focal_length, float img_res, float* camera_translation, float* keypoints_2d,
                                                    void* stream) {
    if (batch < 0 || num_points < 0 || (batch > 0 && (!joints || !pred_camera || !camera_translation || !keypoints_2d)))
        return fail("weak_perspective_projection: bad arguments");
    CUDA_OK(launch_weak_persp_fwd(joints, pred_camera, focal_length, img_res, camera_translation, keypoints_2d, batch, num_points,
                                  static_cast<cudaStream_t>(stream)));
    if (batch) ++g_launches;
    return 0;
}
extern "C" int smplb200_weak_perspective_projection_backward(int batch, int num_points, const float* joints, const float* pred_camera,
                                                             float focal_length, float img_res, const float* grad_keypoints_2d,
                                                             const float* grad_camera_translation, float* grad_joints,
                                                             float* grad_pred_camera, void* stream) {
    if (batch < 0 || num_points < 0 || (batch > 0 && (!joints || !pred_camera || !grad_keypoints_2d || !grad_joints || !grad_pred_camera)))
        return fail("weak_perspective_projection_backward: bad arguments");
    CUDA_OK(launch_weak_persp_bwd(joints, pred_camera, focal_length, img_res, grad_keypoints_2d, grad_camera_translation, grad_joints,
                                  grad_pred_camera, batch, num_points, static_cast<cudaStream_t>(stream)));
    if (batch) ++g_launches;
    return 0;
}

// ---- the steps either side of SMPLify -------------------------------------------------------------------
extern "C" int smplb200_rot6d_to_rotmat(int n, const float* x6, float* rotmat, void* stream) {
    if (n < 0 || (n > 0 && (!x6 || !rotmat))) return fail("rot6d_to_rotmat: bad arguments");
    CUDA_OK(launch_rot6d_to_rotmat(x6, rotmat, n, static_cast<cudaStream_t>(stream)));
    if (n) ++g_launches;
    return 0;
}
extern "C" int smplb200_rotmat_to_axis_angle(int n, const float* rotmat, float* axis_angle, int scrub_nan, void* stream) {
    if (n < 0 || (n > 0 && (!rotmat || !axis_angle))) return fail("rotmat_to_axis_angle: bad arguments");
    CUDA_OK(launch_rotmat_to_aa(rotmat, axis_angle, n, scrub_nan, static_cast<cudaStream_t>(stream)));
    if (n) ++g_launches;
    return 0;
}
extern "C" int smplb200_estimate_translation(int batch, const float* joints3d, const float* keypoints_2d, float focal_length,
                                             float img_size, float* translation, void* stream) {
    if (batch < 0 || (batch > 0 && (!joints3d || !keypoints_2d || !translation))) return fail("estimate_translation: bad arguments");
    CUDA_OK(launch_estimate_translation(joints3d, keypoints_2d, focal_length, img_size, translation, batch, static_cast<cudaStream_t>(stream)));
    if (batch) ++g_launches;
    return 0;
}
extern "C" int smplb200_fits_get(int batch, const float* store, int64_t store_rows, const int64_t* index, const float* rot_deg,
                                 const uint8_t* flipped, const int32_t* pose_flip_perm, float* pose, float* betas, int32_t* status,
                                 void* stream) {
    if (batch < 0) return fail("fits_get: negative batch");
    if (batch == 0) return 0;
    if (!store || store_rows < 0 || !index || !rot_deg || !flipped || !pose_flip_perm || !pose || !betas)
        return fail("fits_get: bad arguments");
    for (int i = 0; i < 72; ++i)
        if (pose_flip_perm[i] < 0 || pose_flip_perm[i] >= 72) return fail("fits_get: pose_flip_perm[%d] out of range", i);
    CUDA_OK(launch_fits_get(store, (long long)store_rows, reinterpret_cast<const long long*>(index), rot_deg, flipped, pose_flip_perm,
                            pose, betas, status, batch, static_cast<cudaStream_t>(stream)));
    ++g_launches;
    return 0;
}
extern "C" int smplb200_fits_set(int batch, float* store, int64_t store_rows, const int64_t* index, const float* rot_deg,
                                 const uint8_t* flipped, const uint8_t* update, const int32_t* pose_flip_perm, const float* pose,
                                 const float* betas, int32_t* status, void* stream) {
    if (batch < 0) return fail("fits_set: negative batch");
    if (batch == 0) return 0;
    if (!store || store_rows < 0 || !index || !rot_deg || !flipped || !update || !pose_flip_perm || !pose || !betas)
        return fail("fits_set: bad arguments");
    for (int i = 0; i < 72; ++i)
        if (pose_flip_perm[i] < 0 || pose_flip_perm[i] >= 72) return fail("fits_set: pose_flip_perm[%d] out of range", i);
    CUDA_OK(launch_fits_set(store, (long long)store_rows, reinterpret_cast<const long long*>(index), rot_deg, flipped, update,
                            pose_flip_perm, pose, betas, status, batch, static_cast<cudaStream_t>(stream)));
    ++g_launches;
    return 0;
}
extern "C" int smplb200_keep_better(int batch, const float* new_reprojection_loss, const float* new_pose, const float* new_betas,
                                    const float* new_cam_t, const float* new_joints, float* best_loss, float* best_pose,
                                    float* best_betas, float* best_cam_t, float* best_joints, uint8_t* update, void* stream) {
    if (batch < 0 || (batch > 0 && (!new_reprojection_loss || !new_pose || !new_betas || !new_cam_t || !best_loss || !best_pose ||
                                    !best_betas || !best_cam_t || !update)))
        return fail("keep_better: bad arguments");
    CUDA_OK(launch_keep_better(new_reprojection_loss, new_pose, new_betas, new_cam_t, new_joints, best_loss, best_pose, best_betas,
                               best_cam_t, best_joints, update, batch, static_cast<cudaStream_t>(stream)));
    if (batch) ++g_launches;
    return 0;
}

// ---- what the train step does with the SMPLify result ---------------------------------------------
extern "C" int smplb200_finalize_fits(int batch, float smplify_threshold, const uint8_t* has_smpl, const float* gt_pose,
                                      const float* gt_betas, const float* gt_cam_t, const float* gt_joints, const float* gt_vertices,
                                      const float* opt_joint_loss, float* opt_pose, float* opt_betas, float* opt_cam_t,
                                      float* opt_joints, float* opt_vertices, uint8_t* valid_fit, void* stream) {
    if (batch < 0 || (batch > 0 && (!has_smpl || !gt_pose || !gt_betas || !gt_cam_t || !gt_joints || !opt_joint_loss || !opt_pose ||
                                    !opt_betas || !opt_cam_t || !opt_joints || !valid_fit || (!opt_vertices != !gt_vertices))))
        return fail("finalize_fits: bad arguments");
    CUDA_OK(launch_finalize_fits(batch, smplify_threshold, has_smpl, gt_pose, gt_betas, gt_cam_t, gt_joints, gt_vertices, opt_joint_loss,
                                 opt_pose, opt_betas, opt_cam_t, opt_joints, opt_vertices, valid_fit, static_cast<cudaStream_t>(stream)));
    if (batch) ++g_launches;
    return 0;
}
extern "C" size_t smplb200_train_loss_workspace_bytes(int batch) { return train_loss_workspace_doubles(batch) * sizeof(double); }
extern "C" int smplb200_smpl_param_losses(int batch, const float* pred_rotmat, const float* pred_betas, const float* gt_pose,
                                          const float* gt_betas, const uint8_t* valid, float* losses, float* grad_pred_rotmat,
                                          float* grad_pred_betas, void* workspace, void* stream) {
    if (batch < 0 || !losses || !workspace || (batch > 0 && (!pred_rotmat || !pred_betas || !gt_pose || !gt_betas || !valid)))
        return fail("smpl_param_losses: bad arguments");
    CUDA_OK(launch_smpl_param_losses(batch, pred_rotmat, pred_betas, gt_pose, gt_betas, valid, losses, grad_pred_rotmat, grad_pred_betas,
                                     static_cast<double*>(workspace), static_cast<cudaStream_t>(stream)));
    g_launches += batch ? 3 : 2;
    return 0;
}
extern "C" int smplb200_keypoint_loss(int batch, const float* pred_keypoints_2d, const float* gt_keypoints_2d, float openpose_weight,
                                      float gt_weight, float* loss, float* grad_pred, void* workspace, void* stream) {
    if (batch < 0 || !loss || !workspace || (batch > 0 && (!pred_keypoints_2d || !gt_keypoints_2d)))
        return fail("keypoint_loss: bad arguments");
    CUDA_OK(launch_keypoint_loss(batch, pred_keypoints_2d, gt_keypoints_2d, openpose_weight, gt_weight, loss, grad_pred,
                                 static_cast<double*>(workspace), static_cast<cudaStream_t>(stream)));
    g_launches += batch ? 3 : 2;
    return 0;
}
extern "C" int smplb200_keypoint_3d_loss(int batch, const float* pred_joints, const float* gt_keypoints_3d, const uint8_t* has_pose_3d,
                                         float* loss, float* grad_pred_joints, void* workspace, void* stream) {
    if (batch < 0 || !loss || !workspace || (batch > 0 && (!pred_joints || !gt_keypoints_3d || !has_pose_3d)))
        return fail("keypoint_3d_loss: bad arguments");
    CUDA_OK(launch_keypoint3d_loss(batch, pred_joints, gt_keypoints_3d, has_pose_3d, loss, grad_pred_joints,
                                   static_cast<double*>(workspace), static_cast<cudaStream_t>(stream)));
    g_launches += batch ? 3 : 2;
    return 0;
}
extern "C" int smplb200_shape_loss(int batch, const float* pred_vertices, const float* gt_vertices, const uint8_t* valid, float* loss,
                                   float* grad_pred_vertices, void* workspace, void* stream) {
    if (batch < 0 || !loss || !workspace || (batch > 0 && (!pred_vertices || !gt_vertices || !valid)))
        return fail("shape_loss: bad arguments");
    CUDA_OK(launch_shape_loss(batch, pred_vertices, gt_vertices, valid, loss, grad_pred_vertices, static_cast<double*>(workspace),
                              static_cast<cudaStream_t>(stream)));
    g_launches += batch ? 3 : 2;
    return 0;
}

// Host-buffer wrapper: H2D of the five inputs, fit, D2H of the results, synchronise.
extern "C" int smplb200_smplify_fit_host(const smplb200_model* cm, int batch, int num_iters, double step_size, float focal_length,
                                         const float* init_pose, const float* init_betas, const float* init_cam_t,
                                         const float* camera_center, float* keypoints_2d, float* vertices, float* joints,
                                         float* pose, float* betas, float* camera_translation, float* reprojection_loss) {
    if (!cm) return fail("NULL model");
    if (batch <= 0) return batch == 0 ? 0 : fail("negative batch");
    if (!init_pose || !init_betas || !init_cam_t || !camera_center || !keypoints_2d)
        return fail("smplify_fit_host: NULL input buffer");
    if (!joints || !pose || !betas || !camera_translation || !reprojection_loss) return fail("smplify_fit_host: NULL output buffer");
    if (!cm->has_prior) return fail("smplify_fit_host: the model was created without a GMM prior");
    if (num_iters < 0) return fail("num_iters must not be negative");
    smplb200_model* m = const_cast<smplb200_model*>(cm);
    std::lock_guard<std::mutex> lock(m->mu);
    DeviceGuard guard(m->device);
    CUDA_OK(guard.err);
    if (!m->host_stream) CUDA_OK(cudaStreamCreateWithFlags(&m->host_stream, cudaStreamNonBlocking));
    if (!m->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    if (!m->fit_done) CUDA_OK(cudaEventCreateWithFlags(&m->fit_done, cudaEventDisableTiming));
    const size_t B = (size_t)batch;
    const size_t n_in = B * (72 + 10 + 3 + 2 + 147), n_out = B * (147 + 72 + 10 + 3 + 49);
    const size_t need = align256(n_in * 4) + align256(n_out * 4) + align256(B * kCols * 4) +
                        smplb200_fit_workspace_bytes(batch) + 1024;
    if (m->scratch_bytes < need) {
        if (m->scratch) CUDA_OK(cudaFree(m->scratch));
        m->scratch = nullptr;
        m->scratch_bytes = 0;
        CUDA_OK(cudaMalloc(&m->scratch, need));
        m->scratch_bytes = need;
    }
    cudaStream_t st = m->host_stream;
    char* base = static_cast<char*>(m->scratch);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_pose = d_in; float* d_betas = d_pose + B * 72; float* d_cam = d_betas + B * 10;
    float* d_cen = d_cam + B * 3; float* d_kp = d_cen + B * 2;
    float* d_out = reinterpret_cast<float*>(base + align256(n_in * 4));
    float* o_joints = d_out; float* o_pose = o_joints + B * 147; float* o_betas = o_pose + B * 72;
    float* o_cam = o_betas + B * 10; float* o_reproj = o_cam + B * 3;
    char* after = base + align256(n_in * 4) + align256(n_out * 4);
    float* d_verts = reinterpret_cast<float*>(after);      // always computed, like the reference; copied back on request
    after += align256(B * kCols * 4);
    void* ws = after;
    CUDA_OK(cudaMemcpyAsync(d_pose, init_pose, B * 72 * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_betas, init_betas, B * 10 * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_cam, init_cam_t, B * 3 * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_cen, camera_center, B * 2 * 4, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemcpyAsync(d_kp, keypoints_2d, B * 147 * 4, cudaMemcpyHostToDevice, st));
    if (run_fit(m, batch, num_iters, step_size, focal_length, 0, d_pose, d_betas, d_cam, d_cen, d_kp, d_verts, o_joints, o_pose,
                o_betas, o_cam, o_reproj, nullptr, ws, smplb200_fit_workspace_bytes(batch), st, m->fit_done))
        return 1;
    // the fit results leave on a second stream as soon as the fit kernel is done, while `st` runs the vertex kernels
    cudaStream_t cs = m->copy_stream;
    CUDA_OK(cudaStreamWaitEvent(cs, m->fit_done, 0));
    CUDA_OK(cudaMemcpyAsync(joints, o_joints, B * 147 * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(pose, o_pose, B * 72 * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(betas, o_betas, B * 10 * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(camera_translation, o_cam, B * 3 * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(reprojection_loss, o_reproj, B * 49 * 4, cudaMemcpyDeviceToHost, cs));
    CUDA_OK(cudaMemcpyAsync(keypoints_2d, d_kp, B * 147 * 4, cudaMemcpyDeviceToHost, cs));   // in-place confidence zeroing
    if (vertices) CUDA_OK(cudaMemcpyAsync(vertices, d_verts, B * kCols * 4, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(cs));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
