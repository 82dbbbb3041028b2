// What the reference's train step does with the SMPLify result, one pass over HBM each:
//
//   finalize_fits_kernel     train/trainer.py:735-748   extreme-betas reset, ground-truth overwrite where has_smpl, valid_fit
//   smpl_param_loss_kernel   train/trainer.py:165-178   smpl_losses: MSE(pred_rotmat, batch_rodrigues(gt_pose)), MSE(betas) on valid rows
//   keypoint_loss_kernel     train/trainer.py:88-98     confidence-weighted squared 2D keypoint error, mean over [B,49,2]
//   keypoint3d_loss_kernel   train/trainer.py:100-117   pelvis-centred, confidence-weighted 3D keypoint error on rows with 3D labels
//   shape_loss_kernel        train/trainer.py:158-164   L1 between predicted and fitted vertices on valid rows
//
// Every loss kernel also writes d(loss)/d(prediction) (optional), so a loss costs one read of its inputs; the mean's
// denominator (the number of selected rows) is counted on the device first - no host synchronisation, unlike the
// reference's boolean-mask indexing.  Reductions are deterministic: fixed-shape tree inside a block, one double partial
// per block, summed in index order by a single block.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "launch.h"

namespace smplb200 {

namespace {

constexpr int kVertFloats = 6890 * 3;
constexpr int kShapeChunks = 5;                       // blocks per vertex row (20670 = 5 x 4134 floats = 5 x 2067 float2)
constexpr int kChunkPairs = kVertFloats / 2 / kShapeChunks;
static_assert(kChunkPairs * 2 * kShapeChunks == kVertFloats, "vertex row must split evenly");

// Block-wide sum in double; result valid in thread 0.  Fixed tree -> bitwise reproducible.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red /*[NT/32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = v;
    __syncthreads();
    double a = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < NT / 32; ++i) a += red[i];
    return a;
}

// ws layout (doubles): [0] reserved, [1] selected-row count (as double), [8 ..] per-block partials
__global__ void __launch_bounds__(1024) mask_count_kernel(const uint8_t* __restrict__ mask, int n, double* __restrict__ ws) {
    __shared__ double red[32];
    double c = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) c += mask ? (mask[i] != 0) : 1;
    const double a = block_sum<1024>(c, red);
    if (threadIdx.x == 0) ws[1] = a;
}

// losses[k] = (sum over blocks of partial[k]) / (count * per_row), 0 when count == 0; nk losses interleaved in partials.
__global__ void __launch_bounds__(256) loss_reduce_kernel(const double* __restrict__ ws, int nblocks, int nk, float per_row0,
                                                          float per_row1, float* __restrict__ losses) {
    __shared__ double red[8];
    const double cnt = ws[1];
    for (int k = 0; k < nk; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += 256) s += ws[8 + (size_t)i * nk + k];
        const double a = block_sum<256>(s, red);
        if (threadIdx.x == 0) losses[k] = cnt > 0.0 ? (float)(a / (cnt * (double)(k == 0 ? per_row0 : per_row1))) : 0.f;
        __syncthreads();
    }
}

// utils/geometry.py:9-45 batch_rodrigues (quaternion form), one joint
__device__ __forceinline__ void quat_rodrigues(const float* th, float* R) {
    const float rx = th[0], ry = th[1], rz = th[2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float L = sqrtf(ax * ax + ay * ay + az * az);
    const float nx = rx / L, ny = ry / L, nz = rz / L;
    float sh, ch;
    sincosf(L * 0.5f, &sh, &ch);
    float w = ch, x = sh * nx, y = sh * ny, z = sh * nz;
    const float N = sqrtf(w * w + x * x + y * y + z * z);
    w /= N; x /= N; y /= N; z /= N;
    const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

// trainer.py:165-178.  One block per sample; partials [b][2] = (sum (pred_R - R(gt_pose))^2, sum (pred_betas - gt_betas)^2).
__global__ void __launch_bounds__(256) smpl_param_loss_kernel(const float* __restrict__ pred_rotmat, const float* __restrict__ pred_betas,
                                                              const float* __restrict__ gt_pose, const float* __restrict__ gt_betas,
                                                              const uint8_t* __restrict__ valid, double* __restrict__ ws,
                                                              float* __restrict__ grad_rotmat, float* __restrict__ grad_betas) {
    __shared__ float R[216];
    __shared__ double red[8];
    const int b = blockIdx.x, t = threadIdx.x;
    const bool on = valid[b] != 0;                       // block-uniform
    const double cnt = ws[1];
    if (on && t < 24) quat_rodrigues(gt_pose + (size_t)b * 72 + 3 * t, R + 9 * t);
    __syncthreads();
    float dp = 0.f, db = 0.f;
    if (t < 216) {
        if (on) dp = pred_rotmat[(size_t)b * 216 + t] - R[t];
        if (grad_rotmat) grad_rotmat[(size_t)b * 216 + t] = on ? (float)(2.0 * dp / (cnt * 216.0)) : 0.f;
    } else if (t < 226) {
        const int l = t - 216;
        if (on) db = pred_betas[(size_t)b * 10 + l] - gt_betas[(size_t)b * 10 + l];
        if (grad_betas) grad_betas[(size_t)b * 10 + l] = on ? (float)(2.0 * db / (cnt * 10.0)) : 0.f;
    }
    const double sp = block_sum<256>((double)dp * dp, red);
    __syncthreads();
    const double sb = block_sum<256>((double)db * db, red);
    if (t == 0) { ws[8 + 2 * (size_t)b] = sp; ws[8 + 2 * (size_t)b + 1] = sb; }
}

// trainer.py:88-98.  conf = gt[..., 2] * (openpose_weight for slots < 25, gt_weight after); mean over all B*49*2 entries.
__global__ void __launch_bounds__(128) keypoint_loss_kernel(const float* __restrict__ pred /*[B][49][2]*/, const float* __restrict__ gt /*[B][49][3]*/,
                                                            float op_w, float gt_w, int batch, double* __restrict__ ws,
                                                            float* __restrict__ grad_pred) {
    __shared__ double red[4];
    const int b = blockIdx.x, t = threadIdx.x;
    float v = 0.f;
    if (t < 98) {
        const int j = t >> 1, c = t & 1;
        const float conf = gt[((size_t)b * 49 + j) * 3 + 2] * (j < 25 ? op_w : gt_w);
        const float d = pred[(size_t)b * 98 + t] - gt[((size_t)b * 49 + j) * 3 + c];
        v = conf * (d * d);
        if (grad_pred) grad_pred[(size_t)b * 98 + t] = (float)(2.0 * (double)conf * d / ((double)batch * 98.0));
    }
    const double s = block_sum<128>((double)v, red);
    if (t == 0) ws[8 + b] = s;
}

// trainer.py:100-117.  pred = joints[:, 25:], both sets centred on the mean of their joints 2 and 3 (the hips).
__global__ void __launch_bounds__(96) keypoint3d_loss_kernel(const float* __restrict__ pred_joints /*[B][49][3]*/,
                                                             const float* __restrict__ gt /*[B][24][4]*/, const uint8_t* __restrict__ has3d,
                                                             double* __restrict__ ws, float* __restrict__ grad_pred /*[B][49][3]*/) {
    __shared__ double red[3];
    __shared__ float gsum[3];                             // sum over joints of conf * 2 d per coordinate (for the pelvis term)
    const int b = blockIdx.x, t = threadIdx.x;
    const bool on = has3d[b] != 0;
    const double cnt = ws[1];
    float v = 0.f, g = 0.f;
    const int j = t / 3, c = t % 3;
    if (t < 3) gsum[t] = 0.f;
    __syncthreads();
    if (on && t < 72) {
        const float* P = pred_joints + ((size_t)b * 49 + 25) * 3;
        const float* G = gt + (size_t)b * 96;
        const float pp = (P[2 * 3 + c] + P[3 * 3 + c]) / 2.f, gp = (G[2 * 4 + c] + G[3 * 4 + c]) / 2.f;
        const float d = (P[j * 3 + c] - pp) - (G[j * 4 + c] - gp);
        const float conf = G[j * 4 + 3];
        v = conf * (d * d);
        g = 2.f * conf * d;
    }
    // fixed-order sum of g over the 24 joints of coordinate c (thread c of the first three)
    __shared__ float gs[72];
    if (t < 72) gs[t] = g;
    __syncthreads();
    if (t < 3) {
        float a = 0.f;
        for (int k = 0; k < 24; ++k) a += gs[3 * k + t];
        gsum[t] = a;
    }
    __syncthreads();
    if (grad_pred) {
        // rows 0..24 of the 49 joints receive no gradient
        for (int i = t; i < 147; i += 96) {
            float o = 0.f;
            if (on && i >= 75) {
                const int jj = (i - 75) / 3, cc = (i - 75) % 3;
                float gg = gs[i - 75];
                if (jj == 2 || jj == 3) gg -= 0.5f * gsum[cc];
                o = (float)((double)gg / (cnt * 72.0));
            }
            grad_pred[(size_t)b * 147 + i] = o;
        }
    }
    const double s = block_sum<96>((double)v, red);
    if (t == 0) ws[8 + b] = s;
}

// trainer.py:158-164 (nn.L1Loss on the rows with valid fits).  kShapeChunks blocks per vertex row, float2 accesses
// (a row is 82 680 bytes: 8-byte aligned for every b).
__global__ void __launch_bounds__(256) shape_loss_kernel(const float2* __restrict__ pred, const float2* __restrict__ gt,
                                                         const uint8_t* __restrict__ valid, double* __restrict__ ws,
                                                         float2* __restrict__ grad_pred) {
    __shared__ double red[8];
    const int b = blockIdx.x / kShapeChunks, ch = blockIdx.x % kShapeChunks, t = threadIdx.x;   // batch in gridDim.x: no 65 535 cap
    const bool on = valid[b] != 0;
    const size_t base = (size_t)b * (kVertFloats / 2) + (size_t)ch * kChunkPairs;
    float acc = 0.f;
    if (on) {
        const float gscale = (float)(1.0 / (ws[1] * (double)kVertFloats));
        for (int i = t; i < kChunkPairs; i += 256) {
            const float2 p = __ldg(pred + base + i), q = __ldg(gt + base + i);
            const float dx = p.x - q.x, dy = p.y - q.y;
            acc += fabsf(dx) + fabsf(dy);
            if (grad_pred) {
                float2 g;
                g.x = dx > 0.f ? gscale : (dx < 0.f ? -gscale : 0.f);
                g.y = dy > 0.f ? gscale : (dy < 0.f ? -gscale : 0.f);
                grad_pred[base + i] = g;
            }
        }
    } else if (grad_pred) {
        for (int i = t; i < kChunkPairs; i += 256) grad_pred[base + i] = make_float2(0.f, 0.f);
    }
    const double s = block_sum<256>((double)acc, red);
    if (t == 0) ws[8 + (size_t)b * kShapeChunks + ch] = s;
}

// trainer.py:735-748.  One block per sample.
__global__ void __launch_bounds__(256) finalize_fits_kernel(float threshold, const uint8_t* __restrict__ has_smpl,
                                                            const float* __restrict__ gt_pose, const float* __restrict__ gt_betas,
                                                            const float* __restrict__ gt_cam, const float* __restrict__ gt_joints,
                                                            const float2* __restrict__ gt_verts, const float* __restrict__ loss,
                                                            float* __restrict__ pose, float* __restrict__ betas, float* __restrict__ cam,
                                                            float* __restrict__ joints, float2* __restrict__ verts,
                                                            uint8_t* __restrict__ valid_fit) {
    __shared__ int s_extreme;
    const int b = blockIdx.x, t = threadIdx.x;
    const bool has = has_smpl[b] != 0;
    if (t == 0) {
        int ex = 0;
        for (int l = 0; l < 10; ++l) ex |= fabsf(betas[(size_t)b * 10 + l]) > 3.f;      // NaN compares false, as in torch
        s_extreme = ex;
        valid_fit[b] = (uint8_t)((loss[b] < threshold) || has);
    }
    __syncthreads();
    if (has) {
        if (t < 72) pose[(size_t)b * 72 + t] = gt_pose[(size_t)b * 72 + t];
        else if (t < 82) betas[(size_t)b * 10 + t - 72] = gt_betas[(size_t)b * 10 + t - 72];
        else if (t < 85) cam[(size_t)b * 3 + t - 82] = gt_cam[(size_t)b * 3 + t - 82];
        if (t < 147) joints[(size_t)b * 147 + t] = gt_joints[(size_t)b * 147 + t];
        if (verts) {
            const size_t base = (size_t)b * (kVertFloats / 2);
            for (int i = t; i < kVertFloats / 2; i += 256) verts[base + i] = __ldg(gt_verts + base + i);
        }
    } else if (s_extreme && t < 10) {
        betas[(size_t)b * 10 + t] = 0.f;
    }
}

}  // namespace

size_t train_loss_workspace_doubles(int batch) { return 8 + (size_t)(batch > 0 ? batch : 0) * kShapeChunks; }

cudaError_t launch_finalize_fits(int batch, float threshold, const uint8_t* has_smpl, const float* gt_pose, const float* gt_betas,
                                 const float* gt_cam, const float* gt_joints, const float* gt_verts, const float* loss, float* pose,
                                 float* betas, float* cam, float* joints, float* verts, uint8_t* valid_fit, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    finalize_fits_kernel<<<batch, 256, 0, st>>>(threshold, has_smpl, gt_pose, gt_betas, gt_cam, gt_joints,
                                                reinterpret_cast<const float2*>(gt_verts), loss, pose, betas, cam, joints,
                                                reinterpret_cast<float2*>(verts), valid_fit);
    return cudaGetLastError();
}

cudaError_t launch_smpl_param_losses(int batch, const float* pred_rotmat, const float* pred_betas, const float* gt_pose,
                                     const float* gt_betas, const uint8_t* valid, float* losses2, float* grad_rotmat, float* grad_betas,
                                     double* ws, cudaStream_t st) {
    mask_count_kernel<<<1, 1024, 0, st>>>(valid, batch, ws);
    if (batch > 0)
        smpl_param_loss_kernel<<<batch, 256, 0, st>>>(pred_rotmat, pred_betas, gt_pose, gt_betas, valid, ws, grad_rotmat, grad_betas);
    loss_reduce_kernel<<<1, 256, 0, st>>>(ws, batch, 2, 216.f, 10.f, losses2);
    return cudaGetLastError();
}

cudaError_t launch_keypoint_loss(int batch, const float* pred, const float* gt, float op_w, float gt_w, float* loss, float* grad_pred,
                                 double* ws, cudaStream_t st) {
    mask_count_kernel<<<1, 1024, 0, st>>>(nullptr, batch, ws);            // every row counts: denominator B * 98
    if (batch > 0) keypoint_loss_kernel<<<batch, 128, 0, st>>>(pred, gt, op_w, gt_w, batch, ws, grad_pred);
    loss_reduce_kernel<<<1, 256, 0, st>>>(ws, batch, 1, 98.f, 98.f, loss);
    return cudaGetLastError();
}

cudaError_t launch_keypoint3d_loss(int batch, const float* pred_joints, const float* gt, const uint8_t* has3d, float* loss,
                                   float* grad_pred, double* ws, cudaStream_t st) {
    mask_count_kernel<<<1, 1024, 0, st>>>(has3d, batch, ws);
    if (batch > 0) keypoint3d_loss_kernel<<<batch, 96, 0, st>>>(pred_joints, gt, has3d, ws, grad_pred);
    loss_reduce_kernel<<<1, 256, 0, st>>>(ws, batch, 1, 72.f, 72.f, loss);
    return cudaGetLastError();
}

cudaError_t launch_shape_loss(int batch, const float* pred, const float* gt, const uint8_t* valid, float* loss, float* grad_pred,
                              double* ws, cudaStream_t st) {
    mask_count_kernel<<<1, 1024, 0, st>>>(valid, batch, ws);
    if (batch > 0)
        shape_loss_kernel<<<(unsigned)batch * kShapeChunks, 256, 0, st>>>(reinterpret_cast<const float2*>(pred),
                                                                     reinterpret_cast<const float2*>(gt), valid, ws,
                                                                     reinterpret_cast<float2*>(grad_pred));
    loss_reduce_kernel<<<1, 256, 0, st>>>(ws, batch * kShapeChunks, 1, (float)kVertFloats, (float)kVertFloats, loss);
    return cudaGetLastError();
}

}  // namespace smplb200
