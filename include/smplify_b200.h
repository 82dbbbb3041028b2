/* C ABI of the B200-native SMPLify / SMPL library (libsmplify_b200.so).
 *
 * The reference (AnonymousSubmission43/Inbed_pose_estimation) has no FFI: its boundary for this
 * path is a set of Python callables.  Each entry point below names the reference callable it
 * replaces (paths into the reference tree); inbed_pose_estimation_b200/ binds them with ctypes
 * underneath Python classes that keep the reference's signatures.
 *
 * Conventions
 *  - every pointer marked "device" is a CUDA device pointer to contiguous row-major fp32 data;
 *    the caller owns all input, output and workspace buffers; the library owns only the
 *    immutable model-constant blob between model_create and model_destroy;
 *  - every compute call is asynchronous on `stream` (a cudaStream_t passed as void*) and is
 *    thread-compatible; the only process-wide state is a per-device cache of the SM count.  Compute
 *    calls run on the CALLER's current device (the one the model was created on); model_create,
 *    model_destroy and smplify_fit_host switch to the model's device themselves and restore the
 *    caller's current device before returning;
 *  - return value 0 = success; anything else is an error whose text smplb200_last_error()
 *    returns for the calling thread.  There is no CPU fallback: without a CUDA device
 *    model_create fails.
 */
#ifndef SMPLIFY_B200_H_
#define SMPLIFY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPLB200_NUM_VERTS 6890
#define SMPLB200_NUM_JOINTS 24
#define SMPLB200_NUM_BETAS 10
#define SMPLB200_NUM_POSE_FEATURES 207
#define SMPLB200_NUM_OUT_JOINTS 49
#define SMPLB200_NUM_GAUSSIANS 8
#define SMPLB200_MAX_ITERS 256 /* Adam scalars of the first 256 steps per stage are tabulated on chip; longer fits work, a little slower */
#define SMPLB200_VPOSED_PITCH 20736 /* floats per sample of the saved v_posed buffer (3*6890 padded; rows 16-byte aligned) */

typedef struct smplb200_model smplb200_model;

/* Host-side description of the model constants (all host pointers, copied during create).
 * Mirrors what models/smpl.py:14-19 + smplx.SMPL.__init__ register as buffers and what
 * smplify/prior.py:142-160 derives from gmm_08.pkl. */
typedef struct {
    const float* v_template;         /* [6890][3] */
    const float* shapedirs;          /* [6890][3][10] */
    const float* posedirs;           /* [6890][3][207]  (file layout, before smplx's reshape/transpose) */
    const float* J_regressor;        /* [24][6890] */
    const float* weights;            /* [6890][24] */
    const float* J_regressor_extra;  /* [9][6890]   models/smpl.py:17-18 */
    const int32_t* parents;          /* [24], root = -1 */
    const int32_t* extra_vertex_ids; /* [21]  smplx VertexJointSelector */
    const int32_t* joint_map;        /* [49] -> index into the 54 joints; models/smpl.py:16,19 */
    const int32_t* ign_joints;       /* output joints zeroed for the body stage; smplify/smplify.py:28-29 */
    int32_t num_ign_joints;          /* <= 8 */
    const int32_t* cam_op_joints;    /* [4] smplify/losses.py:72-73 */
    const int32_t* cam_gt_joints;    /* [4] smplify/losses.py:74-75 */
    const int32_t* angle_prior_ids;  /* [4] body_pose entries; smplify/losses.py:24 */
    const float* angle_prior_signs;  /* [4] */
    const float* gmm_means;          /* [8][69]      may be NULL: SMPL only, no fitting */
    const float* gmm_precisions;     /* [8][69][69]  prior.py:146-150 */
    const float* gmm_nll_weights;    /* [8]          prior.py:153-160 */
} smplb200_model_desc;

int smplb200_version(void);
const char* smplb200_last_error(void);

/* Builds the constant blob on CUDA device `device` (folds the joint regressors in float64 on the
 * host, uploads ~60 MB).  Replaces SMPL.__init__ (models/smpl.py:14-19) and
 * MaxMixturePrior.__init__ (smplify/prior.py:102-174) as far as device state is concerned. */
int smplb200_model_create(const smplb200_model_desc* desc, int device, smplb200_model** out);
void smplb200_model_destroy(smplb200_model* model);

/* Bytes of device workspace the calls below need for `batch` samples. */
size_t smplb200_fit_workspace_bytes(int batch);
size_t smplb200_smpl_workspace_bytes(int batch);

/* SMPLify.__call__ (smplify/smplify.py:40-136): two-stage fit, num_iters Adam steps per stage.
 * keypoints [B][49][3] is read AND its confidences at ign_joints are zeroed in place after the
 * camera stage, exactly like the reference (:105).  vertices / loss_trace / packed_results may be NULL.
 * packed_results [B][134]: the fitted pose (72), betas (10), camera translation (3) and per-joint reprojection loss (49)
 * of every sample as one row - the row a sharded refit all-gathers (the kernel writes it besides the separate outputs).
 * loss_trace [2*num_iters][B]: per-sample loss of every iteration (stage 1 then stage 2).
 * step_size is the Adam lr as a double: torch.optim.Adam divides the Python float by the bias
 * correction in float64 before the fp32 cast (smplify.py:79,107), and so does the kernel. */
int smplb200_smplify_fit(const smplb200_model* model, int batch, int num_iters, double step_size, float focal_length,
                         const float* init_pose /*device [B][72]*/, const float* init_betas /*device [B][10]*/,
                         const float* init_cam_t /*device [B][3]*/, const float* camera_center /*device [B][2]*/,
                         float* keypoints_2d /*device [B][49][3], in/out*/,
                         float* vertices /*device [B][6890][3] or NULL*/, float* joints /*device [B][49][3]*/,
                         float* pose /*device [B][72]*/, float* betas /*device [B][10]*/, float* camera_translation /*device [B][3]*/,
                         float* reprojection_loss /*device [B][49]*/, float* loss_trace /*device or NULL*/,
                         float* packed_results /*device [B][134] or NULL*/,
                         void* workspace, size_t workspace_bytes, void* stream);

/* SMPLify.get_fitting_loss (smplify/smplify.py:138-172): zeroes the ignored confidences in place
 * (:156), one forward, per-joint reprojection loss [B][49]. */
int smplb200_smplify_fitting_loss(const smplb200_model* model, int batch, float focal_length,
                                  const float* pose, const float* betas, const float* cam_t, const float* camera_center,
                                  float* keypoints_2d, float* reprojection_loss,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* The three prior terms of body_fitting_loss (smplify/losses.py:46-52) on their own, through the same device code the
 * fit runs every iteration: terms[b] = { 4.78^2 * MaxMixturePrior(body_pose) (smplify/prior.py:181-196, merged
 * max-mixture), 15.2^2 * sum angle_prior(body_pose) (losses.py:19-24), 5^2 * |betas|^2 }.  pose is the full [B][72]
 * pose (entries 3..71 are the body pose).  Optional outputs (may be NULL): the 8 per-component values
 * 0.5 d^T P d - log(nll_weight), the selected component, and the gradient of the sum of the three terms. */
int smplb200_prior_terms(const smplb200_model* model, int batch, const float* pose /*[B][72]*/, const float* betas /*[B][10]*/,
                         float* terms /*[B][3]*/, float* components /*[B][8]*/, int32_t* argmin /*[B]*/,
                         float* grad_body_pose /*[B][69]*/, float* grad_betas /*[B][10]*/, void* stream);

/* SMPL.forward (models/smpl.py:21-33 -> smplx lbs).  rotmat_mode 0: pose is axis-angle [B][72]
 * (global_orient ++ body_pose); 1: pose is [B][24][3][3] (pose2rot=False).  saved_vposed
 * ([B][SMPLB200_VPOSED_PITCH] or NULL) keeps v_posed (what autograd would save) for
 * smplb200_smpl_backward; NULL when no vertex gradient will be requested. */
int smplb200_smpl_forward(const smplb200_model* model, int batch, int rotmat_mode,
                          const float* pose, const float* betas,
                          float* vertices /*[B][6890][3] or NULL*/, float* joints /*[B][49][3]*/,
                          float* saved_vposed /*[B][SMPLB200_VPOSED_PITCH] or NULL*/,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of the above w.r.t. pose (same layout as the input) and betas, given upstream
 * gradients of vertices and/or joints (either may be NULL).  What torch autograd does through
 * smplx in the reference (train/trainer.py:597-615 -> loss.backward()). */
int smplb200_smpl_backward(const smplb200_model* model, int batch, int rotmat_mode,
                           const float* pose, const float* betas, const float* saved_vposed,
                           const float* grad_vertices, const float* grad_joints,
                           float* grad_pose, float* grad_betas,
                           void* workspace, size_t workspace_bytes, void* stream);

/* utils/geometry.py:9-45 batch_rodrigues (quaternion form) and its gradient. */
int smplb200_batch_rodrigues(int n, const float* theta /*[n][3]*/, float* rotmat /*[n][3][3]*/, void* stream);
int smplb200_batch_rodrigues_backward(int n, const float* theta, const float* grad_rotmat, float* grad_theta, void* stream);

/* utils/geometry.py:79-114 perspective_projection and its gradient.  focal_length is a device
 * pointer to 1 value (focal_per_batch = 0) or to [B] values (focal_per_batch = 1).  out_3d != 0
 * is the reference's out_3d=True variant (:108-114; callers train/trainer.py:621-626,
 * models/hmr.py:1720, eval.py:255): projected / grad_projected are then [B][N][3] with the
 * camera-space depth of the point in the third channel. */
int smplb200_perspective_projection(int batch, int num_points, const float* points, const float* rotation,
                                    const float* translation, const float* focal_length, int focal_per_batch,
                                    const float* camera_center, int out_3d, float* projected /*[B][N][2 or 3]*/, void* stream);
int smplb200_perspective_projection_backward(int batch, int num_points, const float* points, const float* rotation,
                                             const float* translation, const float* focal_length, int focal_per_batch,
                                             int out_3d, const float* grad_projected, float* grad_points, float* grad_rotation,
                                             float* grad_translation, void* stream);

/* train/trainer.py:187-199 (and :603-615, models/hmr.py:1708-1710, eval.py:245-247): weak-perspective camera
 * pred_camera [B][3] = (s, tx, ty) -> camera_translation [B][3] = (tx, ty, 2 f / (img_res s + 1e-9)) and the joints
 * [B][N][3] projected with it (identity rotation, zero camera centre), divided by img_res / 2 -> keypoints_2d [B][N][2].
 * The backward call takes d/d keypoints_2d (and optionally d/d camera_translation, may be NULL). */
int smplb200_weak_perspective_projection(int batch, int num_points, const float* joints, const float* pred_camera, float focal_length,
                                         float img_res, float* camera_translation, float* keypoints_2d, void* stream);
int smplb200_weak_perspective_projection_backward(int batch, int num_points, const float* joints, const float* pred_camera,
                                                  float focal_length, float img_res, const float* grad_keypoints_2d,
                                                  const float* grad_camera_translation, float* grad_joints, float* grad_pred_camera,
                                                  void* stream);

/* ---- the steps either side of SMPLify in the reference's train step ------------------------------- */

/* utils/geometry.py:47-61 rot6d_to_rotmat: [n][6] (viewed [n][3][2]) -> [n][3][3]. */
int smplb200_rot6d_to_rotmat(int n, const float* x6, float* rotmat, void* stream);

/* train/trainer.py:702-706: torchgeometry.rotation_matrix_to_angle_axis on [n][3][3] rotation matrices
 * (the homogeneous column the trainer appends is not read) and, with scrub_nan != 0, the NaN -> 0 patch. */
int smplb200_rotmat_to_axis_angle(int n, const float* rotmat, float* axis_angle /*[n][3]*/, int scrub_nan, void* stream);

/* utils/geometry.py:118-181 estimate_translation: per sample weighted least squares over the 24 ground-truth
 * joint slots 25..48 of joints3d [B][49][3] and keypoints_2d [B][49][3] (x, y, conf), float64 inside. */
int smplb200_estimate_translation(int batch, const float* joints3d, const float* keypoints_2d, float focal_length,
                                  float img_size, float* translation /*[B][3]*/, void* stream);

/* train/fits_dict.py:34-94 FitsDict.__getitem__ / __setitem__ on a device-resident store [N][82]
 * (72 pose + 10 betas per image): gather + rotate + flip, and un-flip + un-rotate + masked scatter.
 * index int64 [B], rot_deg fp32 [B], flipped / update uint8 [B], pose_flip_perm = the 72 entries of
 * constants.SMPL_POSE_FLIP_PERM (host pointer).  Rows of one batch must have distinct indices.
 * Indices are validated against store_rows INSIDE the kernel (no host synchronisation): an out-of-range row is
 * returned as NaNs (get) or skipped (set) and bit 0 of *status (device int32, may be NULL) is raised. */
int smplb200_fits_get(int batch, const float* store, int64_t store_rows, const int64_t* index, const float* rot_deg,
                      const uint8_t* flipped, const int32_t* pose_flip_perm, float* pose /*[B][72]*/, float* betas /*[B][10]*/,
                      int32_t* status, void* stream);
int smplb200_fits_set(int batch, float* store, int64_t store_rows, const int64_t* index, const float* rot_deg,
                      const uint8_t* flipped, const uint8_t* update, const int32_t* pose_flip_perm, const float* pose,
                      const float* betas, int32_t* status, void* stream);

/* train/trainer.py:716-727: update[b] = mean_j(new_reprojection_loss[b][j]) < best_loss[b]; where it holds the
 * best_* rows are overwritten by the new fit (best_joints / new_joints [B][49][3] may both be NULL). */
int smplb200_keep_better(int batch, const float* new_reprojection_loss /*[B][49]*/, const float* new_pose, const float* new_betas,
                         const float* new_cam_t, const float* new_joints, float* best_loss /*[B]*/, float* best_pose,
                         float* best_betas, float* best_cam_t, float* best_joints, uint8_t* update /*[B]*/, void* stream);

/* ---- what the train step does with the SMPLify result (train/trainer.py:735-772) ------------------- */

/* train/trainer.py:735-748: opt_betas rows with any |beta| > 3 are zeroed; rows with has_smpl != 0 take the ground-truth
 * pose / betas / cam_t / model joints [B][49][3] / vertices [B][6890][3] (opt_vertices and gt_vertices may both be NULL);
 * valid_fit[b] = (opt_joint_loss[b] < smplify_threshold) | has_smpl[b].  All opt_* are updated in place. */
int smplb200_finalize_fits(int batch, float smplify_threshold, const uint8_t* has_smpl, const float* gt_pose, const float* gt_betas,
                           const float* gt_cam_t, const float* gt_joints, const float* gt_vertices, const float* opt_joint_loss,
                           float* opt_pose, float* opt_betas, float* opt_cam_t, float* opt_joints, float* opt_vertices,
                           uint8_t* valid_fit /*[B]*/, void* stream);

/* Device scratch (bytes) of the four loss entry points below for a batch. */
size_t smplb200_train_loss_workspace_bytes(int batch);

/* Each loss call writes its scalar(s) to DEVICE memory and, when the grad pointer is not NULL, d(loss)/d(prediction)
 * for an upstream gradient of 1 (rows outside the mask get zeros).  An empty mask gives loss 0 and zero gradients
 * (the reference returns a zero tensor in that case).  No host synchronisation.
 *
 * train/trainer.py:165-178 smpl_losses: losses[0] = MSE(pred_rotmat[valid], batch_rodrigues(gt_pose)[valid]),
 * losses[1] = MSE(pred_betas[valid], gt_betas[valid]); pred_rotmat [B][24][3][3], gt_pose [B][72]. */
int smplb200_smpl_param_losses(int batch, const float* pred_rotmat, const float* pred_betas, const float* gt_pose,
                               const float* gt_betas, const uint8_t* valid, float* losses /*[2]*/, float* grad_pred_rotmat,
                               float* grad_pred_betas, void* workspace, void* stream);
/* train/trainer.py:88-98 keypoint_loss: mean over [B][49][2] of conf * (pred - gt)^2, conf = gt[..,2] scaled by
 * openpose_weight (slots 0..24) / gt_weight (slots 25..48). */
int smplb200_keypoint_loss(int batch, const float* pred_keypoints_2d /*[B][49][2]*/, const float* gt_keypoints_2d /*[B][49][3]*/,
                           float openpose_weight, float gt_weight, float* loss /*[1]*/, float* grad_pred, void* workspace,
                           void* stream);
/* train/trainer.py:100-117 keypoint_3d_loss: pred_joints [B][49][3] (slots 25..48 are read), gt_keypoints_3d [B][24][4]
 * (x, y, z, conf), rows with has_pose_3d != 0, both pelvis-centred (mean of joints 2 and 3). */
int smplb200_keypoint_3d_loss(int batch, const float* pred_joints, const float* gt_keypoints_3d, const uint8_t* has_pose_3d,
                              float* loss /*[1]*/, float* grad_pred_joints /*[B][49][3]*/, void* workspace, void* stream);
/* train/trainer.py:158-164 shape_loss: L1 mean over the [6890][3] vertices of the rows with valid != 0. */
int smplb200_shape_loss(int batch, const float* pred_vertices, const float* gt_vertices, const uint8_t* valid, float* loss /*[1]*/,
                        float* grad_pred_vertices, void* workspace, void* stream);

/* Host-buffer convenience wrapper of smplb200_smplify_fit: all pointers are HOST pointers
 * (pinned for best throughput); copies inputs to the device, runs the fit, copies the results
 * back and synchronises.  vertices may be NULL (they are then left on the device and not
 * copied).  This is the call timed as "e2e" by bench.py. */
int smplb200_smplify_fit_host(const smplb200_model* model, int batch, int num_iters, double step_size, float focal_length,
                              const float* init_pose, const float* init_betas, const float* init_cam_t,
                              const float* camera_center, float* keypoints_2d,
                              float* vertices, float* joints, float* pose, float* betas, float* camera_translation,
                              float* reprojection_loss);

/* How smplb200_smplify_fit tiles a batch over `sms` SMs (one CTA per tile, one tile per SM and wave): n16 tiles of 16
 * samples followed by n_small tiles of `small` (4, 8 or 12; 0 = none) samples.  Pure host arithmetic. */
void smplb200_fit_tile_plan(int batch, int sms, int* n16, int* small, int* n_small);

/* Batches of 1024 samples and more run on the PAIR kernel: 2-CTA clusters of 2 x 16 (or, in the last wave, 2 x 12) samples
 * that share the per-iteration GEMMs on the tensor cores (tcgen05 cta_group::2, 3xTF32).  Returns 1 if smplb200_smplify_fit
 * uses it for `batch` (then n16 / n12 = the numbers of 2 x 16- and 2 x 12-sample pairs on `sms` SMs), else 0 (the tile plan
 * above applies).  Pure host arithmetic. */
int smplb200_fit_pair_plan(int batch, int sms, int* n16, int* n12);
/* Small batches (the reference trains with --batch_size 32, README.md:33-35; SMPLify call at train/trainer.py:709-715): a
 * cluster of 8, 4 or 2 CTAs fits each 4-sample tile, the per-iteration GEMMs split by output rows over the cluster
 * (csrc/fit_split.cuh).  Returns the cluster size for `batch`, 0 when the batch is too large for the cluster kernel.
 * sms <= 0: the plan smplb200_smplify_fit uses on the CURRENT device - the largest cluster size whose clusters are all
 * resident at once (cudaOccupancyMaxActiveClusters: clusters live inside one GPC, so fewer than sms / C fit); needs a GPU.
 * sms > 0: the same arithmetic with sms / C clusters assumed resident (an upper bound; no GPU needed). */
int smplb200_fit_split_plan(int batch, int sms);

/* Number of this library's kernel launches issued by the process (any thread: torch autograd runs backward calls on
 * its own threads) since the last reset; reset != 0 returns the count and clears it (bench.py reports it as gpu_launches). */
long long smplb200_launch_count(int reset);

/* Measures the fp32 FMA rate of the CUDA-core pipes on the current device (packed = 0: scalar FFMA,
 * 1: Blackwell packed FFMA2) - the roofline denominator bench.py uses for the SIMT-bound fit kernel. */
int smplb200_probe_fp32_peak(int packed, double* tflops);

/* Measures the dense tcgen05 kind::tf32 rate (M=128 N=256 K=8 MMAs issued back to back on every SM, fp32 accumulators in
 * TMEM) on the current device: the tensor-pipe denominator of the 3xTF32 LBS kernels (one fp32-accurate product costs three
 * such MMAs). */
int smplb200_probe_tf32_peak(double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* SMPLIFY_B200_H_ */
