// Host-callable launchers implemented in kernels.cu (and, later, the tcgen05 kernels).
#pragma once
#include <cuda_runtime.h>
#include "smpl_common.h"

namespace smplb200 {

struct FitParams;
struct PoseParams;
struct PriorParams;

constexpr int kFitThreads = 384;
constexpr int kPoseThreads = 384;

cudaError_t launch_fit(const ModelView& M, const FitParams& P, cudaStream_t stream);
void plan_fit_tiles(int batch, int sms, int* n16, int* small, int* n_small);
void plan_fit_pairs(int batch, int sms, int* n16, int* n12);
int plan_fit_split(int batch, int sms);      // CTAs per 4-sample tile of the small-batch cluster kernel (8 / 4 / 2), 0 = not used: by SM count (upper bound)
int plan_fit_split_device(int batch);        // ... by the cluster occupancy of the current device (what the launch uses)
int fit_uses_pairs(int batch, int num_iters);      // 1 when launch_fit runs the batch on the pair kernel (tensor-core GEMMs)
int device_sm_count();      // of the current device (cached per device index)
cudaError_t launch_prior_terms(const ModelView& M, const PriorParams& P, cudaStream_t stream);
cudaError_t launch_pose_forward(const ModelView& M, const PoseParams& P, cudaStream_t stream);
cudaError_t launch_pose_backward(const ModelView& M, const PoseParams& P, cudaStream_t stream);
cudaError_t launch_quat_rodrigues_fwd(const float* theta, float* rot, int n, cudaStream_t st);
cudaError_t launch_quat_rodrigues_bwd(const float* theta, const float* grot, float* gtheta, int n, cudaStream_t st);
cudaError_t launch_projection_fwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* cen, float* out, int out_3d, int batch, int npts, cudaStream_t st);
cudaError_t launch_projection_bwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* gout, int out_3d, float* gpts, float* grot, float* gtr, int batch, int npts,
                                  cudaStream_t st);

// adjacent.cu: the steps either side of SMPLify in the reference's train step
cudaError_t launch_rot6d_to_rotmat(const float* x, float* R, int n, cudaStream_t st);
cudaError_t launch_rotmat_to_aa(const float* R, float* aa, int n, int scrub_nan, cudaStream_t st);
cudaError_t launch_estimate_translation(const float* S, const float* kp, float focal, float img_size, float* trans, int batch,
                                        cudaStream_t st);
cudaError_t launch_fits_get(const float* store, long long store_rows, const long long* index, const float* rot, const uint8_t* flipped,
                            const int* perm72, float* pose, float* betas, int* status, int batch, cudaStream_t st);
cudaError_t launch_fits_set(float* store, long long store_rows, const long long* index, const float* rot, const uint8_t* flipped,
                            const uint8_t* update, const int* perm72, const float* pose, const float* betas, int* status, int batch,
                            cudaStream_t st);
cudaError_t launch_keep_better(const float* new_reproj, const float* new_pose, const float* new_betas, const float* new_cam,
                               const float* new_joints, float* loss, float* pose, float* betas, float* cam, float* joints,
                               uint8_t* update, int batch, cudaStream_t st);

cudaError_t launch_weak_persp_fwd(const float* joints, const float* cam, float focal, float img_res, float* cam_t, float* kp, int batch,
                                  int npts, cudaStream_t st);
cudaError_t launch_weak_persp_bwd(const float* joints, const float* cam, float focal, float img_res, const float* g_kp,
                                  const float* g_cam_t, float* g_joints, float* g_cam, int batch, int npts, cudaStream_t st);

// train_losses.cu: what the train step does with the SMPLify result
size_t train_loss_workspace_doubles(int batch);
cudaError_t launch_finalize_fits(int batch, float threshold, const uint8_t* has_smpl, const float* gt_pose, const float* gt_betas,
                                 const float* gt_cam, const float* gt_joints, const float* gt_verts, const float* loss, float* pose,
                                 float* betas, float* cam, float* joints, float* verts, uint8_t* valid_fit, cudaStream_t st);
cudaError_t launch_smpl_param_losses(int batch, const float* pred_rotmat, const float* pred_betas, const float* gt_pose,
                                     const float* gt_betas, const uint8_t* valid, float* losses2, float* grad_rotmat, float* grad_betas,
                                     double* ws, cudaStream_t st);
cudaError_t launch_keypoint_loss(int batch, const float* pred, const float* gt, float op_w, float gt_w, float* loss, float* grad_pred,
                                 double* ws, cudaStream_t st);
cudaError_t launch_keypoint3d_loss(int batch, const float* pred_joints, const float* gt, const uint8_t* has3d, float* loss,
                                   float* grad_pred, double* ws, cudaStream_t st);
cudaError_t launch_shape_loss(int batch, const float* pred, const float* gt, const uint8_t* valid, float* loss, float* grad_pred,
                              double* ws, cudaStream_t st);

}  // namespace smplb200
