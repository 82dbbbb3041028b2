// Host-callable launchers implemented in kernels.cu (and, later, the tcgen05 kernels).
#pragma once
#include <cuda_runtime.h>
#include "smpl_common.h"

namespace smplb200 {

struct FitParams;
struct PoseParams;

constexpr int kFitThreads = 384;
constexpr int kPoseThreads = 384;

cudaError_t launch_fit(const ModelView& M, const FitParams& P, cudaStream_t stream);
cudaError_t launch_pose_forward(const ModelView& M, const PoseParams& P, cudaStream_t stream);
cudaError_t launch_pose_backward(const ModelView& M, const PoseParams& P, cudaStream_t stream);
cudaError_t launch_quat_rodrigues_fwd(const float* theta, float* rot, int n, cudaStream_t st);
cudaError_t launch_quat_rodrigues_bwd(const float* theta, const float* grot, float* gtheta, int n, cudaStream_t st);
cudaError_t launch_projection_fwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* cen, float* out, int batch, int npts, cudaStream_t st);
cudaError_t launch_projection_bwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* gout, float* gpts, float* grot, float* gtr, int batch, int npts, cudaStream_t st);

}  // namespace smplb200
