// The steps either side of SMPLify in the reference's train step, one fused kernel each:
//
//   rot6d_to_rotmat_kernel        utils/geometry.py:47-61            CNN output -> rotation matrices
//   rotmat_to_aa_kernel           train/trainer.py:702-706           torchgeometry.rotation_matrix_to_angle_axis + the NaN scrub
//   estimate_translation_kernel   utils/geometry.py:118-181          weighted least squares per sample (float64 like numpy)
//   fits_get_kernel / fits_set_kernel   train/fits_dict.py:34-94     best-fit store gather / masked scatter with the
//                                                                    rotate (tgm aa->R, in-plane rotation, cv2.Rodrigues R->aa)
//                                                                    and flip (constants.SMPL_POSE_FLIP_PERM) transforms
//   keep_better_kernel            train/trainer.py:716-727           update = new_loss < old_loss; masked overwrite
//
// torchgeometry (requirements.txt:13, unpinned, 0.1.x API) and OpenCV are third-party code that is not in the reference
// tree; their published algorithms are restated here (and in oracle/tgm_shim.py for the checker).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "launch.h"

namespace smplb200 {

// ---------------------------------------------------------------------------------------------------------------------
// utils/geometry.py:47-61.  x.view(-1, 3, 2): a1 = x[:, :, 0], a2 = x[:, :, 1]; F.normalize(v) = v / max(|v|, 1e-12).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void rot6d_to_rotmat_kernel(const float* __restrict__ x, float* __restrict__ R, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = x + 6 * (size_t)i;
    const float a1[3] = {p[0], p[2], p[4]}, a2[3] = {p[1], p[3], p[5]};
    const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    const float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    const float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    const float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
    float* o = R + 9 * (size_t)i;                       // torch.stack((b1, b2, b3), dim=-1): columns
#pragma unroll
    for (int r = 0; r < 3; ++r) { o[3 * r + 0] = b1[r]; o[3 * r + 1] = b2[r]; o[3 * r + 2] = b3[r]; }
}

// ---------------------------------------------------------------------------------------------------------------------
// torchgeometry 0.1.x rotation_matrix_to_angle_axis = rotation_matrix_to_quaternion (eps = 1e-6, four-branch form on the
// TRANSPOSED matrix) followed by quaternion_to_angle_axis; trainer.py:706 then replaces NaNs by 0.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tgm_rotmat_to_aa(const float* R, float* aa) {
    // rmat_t[i][j] = R[j][i]
    const float t00 = R[0], t01 = R[3], t02 = R[6], t10 = R[1], t11 = R[4], t12 = R[7], t20 = R[2], t21 = R[5], t22 = R[8];
    const bool d2 = t22 < 1e-6f, d0_d1 = t00 > t11, d0_nd1 = t00 < -t11;
    float q[4], t;
    if (d2 && d0_d1) {
        t = 1.f + t00 - t11 - t22;
        q[0] = t12 - t21; q[1] = t; q[2] = t01 + t10; q[3] = t20 + t02;
    } else if (d2) {
        t = 1.f - t00 + t11 - t22;
        q[0] = t20 - t02; q[1] = t01 + t10; q[2] = t; q[3] = t12 + t21;
    } else if (d0_nd1) {
        t = 1.f - t00 - t11 + t22;
        q[0] = t01 - t10; q[1] = t20 + t02; q[2] = t12 + t21; q[3] = t;
    } else {
        t = 1.f + t00 + t11 + t22;
        q[0] = t; q[1] = t12 - t21; q[2] = t20 - t02; q[3] = t01 - t10;
    }
    const float s = sqrtf(t);
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = (q[k] / s) * 0.5f;
    const float sin2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const float sn = sqrtf(sin2), cs = q[0];
    const float two_theta = 2.0f * ((cs < 0.0f) ? atan2f(-sn, -cs) : atan2f(sn, cs));
    const float k = (sin2 > 0.0f) ? two_theta / sn : 2.0f;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
}

__global__ void rotmat_to_aa_kernel(const float* __restrict__ R, float* __restrict__ aa, int n, int scrub_nan) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[9], o[3];
#pragma unroll
    for (int e = 0; e < 9; ++e) r[e] = R[9 * (size_t)i + e];
    tgm_rotmat_to_aa(r, o);
#pragma unroll
    for (int a = 0; a < 3; ++a) aa[3 * (size_t)i + a] = (scrub_nan && isnan(o[a])) ? 0.f : o[a];
}

// ---------------------------------------------------------------------------------------------------------------------
// utils/geometry.py:118-181.  Per sample: the 24 ground-truth joints (slots 25..48), weights sqrt(conf), normal
// equations of [F 0 cx-u; 0 F cy-v] t = (uv - c) z - F xy in float64 (numpy promotes to float64), solved by Gaussian
// elimination with partial pivoting (what LAPACK gesv does for np.linalg.solve).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void estimate_translation_kernel(const float* __restrict__ S, const float* __restrict__ kp, float focal, float img_size,
                                            float* __restrict__ trans, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double F = (double)focal, c0 = (double)img_size / 2.0;
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rhs[3] = {0, 0, 0};
    for (int j = 25; j < 49; ++j) {
        const float* s = S + ((size_t)b * 49 + j) * 3;
        const float* k = kp + ((size_t)b * 49 + j) * 3;
        const double w = sqrt((double)k[2]);                 // weight2 = sqrt(conf), applied to Q and c (so squared in A, b)
        const double z = (double)s[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double uv = (double)k[a];
            double q[3] = {a == 0 ? F : 0.0, a == 1 ? F : 0.0, c0 - uv};
            double c = (uv - c0) * z - F * (double)s[a];
            q[0] *= w; q[1] *= w; q[2] *= w; c *= w;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) A[r][cc] += q[r] * q[cc];
                rhs[r] += q[r] * c;
            }
        }
    }
    // LU with partial pivoting
    int piv[3] = {0, 1, 2};
    for (int col = 0; col < 3; ++col) {
        int best = col;
        for (int r = col + 1; r < 3; ++r)
            if (fabs(A[piv[r]][col]) > fabs(A[piv[best]][col])) best = r;
        const int tmp = piv[col]; piv[col] = piv[best]; piv[best] = tmp;
        const int pr = piv[col];
        for (int r = col + 1; r < 3; ++r) {
            const int rr = piv[r];
            const double f = A[rr][col] / A[pr][col];
            for (int cc = col; cc < 3; ++cc) A[rr][cc] -= f * A[pr][cc];
            rhs[rr] -= f * rhs[pr];
        }
    }
    double x[3];
    for (int r = 2; r >= 0; --r) {
        const int pr = piv[r];
        double a = rhs[pr];
        for (int cc = r + 1; cc < 3; ++cc) a -= A[pr][cc] * x[cc];
        x[r] = a / A[pr][r];
    }
    trans[3 * (size_t)b + 0] = (float)x[0];
    trans[3 * (size_t)b + 1] = (float)x[1];
    trans[3 * (size_t)b + 2] = (float)x[2];
}

// ---------------------------------------------------------------------------------------------------------------------
// FitsDict transforms (train/fits_dict.py:62-94)
// ---------------------------------------------------------------------------------------------------------------------
// torchgeometry angle_axis_to_rotation_matrix (fp32): theta2 > 1e-6 -> Rodrigues with axis = aa / (theta + 1e-6),
// else first-order Taylor form.
__device__ __forceinline__ void tgm_aa_to_rotmat(const float* aa, float* R) {
    const float theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (theta2 > 1e-6f) {
        const float theta = sqrtf(theta2);
        const float wx = aa[0] / (theta + 1e-6f), wy = aa[1] / (theta + 1e-6f), wz = aa[2] / (theta + 1e-6f);
        const float c = cosf(theta), s = sinf(theta), k = 1.0f - c;
        R[0] = c + wx * wx * k;       R[1] = wx * wy * k - wz * s;  R[2] = wy * s + wx * wz * k;
        R[3] = wz * s + wx * wy * k;  R[4] = c + wy * wy * k;       R[5] = -wx * s + wy * wz * k;
        R[6] = -wy * s + wx * wz * k; R[7] = wx * s + wy * wz * k;  R[8] = c + wz * wz * k;
    } else {
        R[0] = 1.f; R[1] = -aa[2]; R[2] = aa[1];
        R[3] = aa[2]; R[4] = 1.f; R[5] = -aa[0];
        R[6] = -aa[1]; R[7] = aa[0]; R[8] = 1.f;
    }
}

// cv2.Rodrigues, matrix -> vector, in float64 (OpenCV first projects the matrix onto SO(3) with an SVD; the inputs here
// are products of two rotations, orthonormal to fp32 rounding, for which that projection changes nothing above 1e-7).
__device__ __forceinline__ void cv_rotmat_to_aa(const double* R, double* r) {
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    const double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
        double t = (R[0] + 1.0) * 0.5;
        rx = sqrt(t > 0.0 ? t : 0.0);
        t = (R[4] + 1.0) * 0.5;
        ry = sqrt(t > 0.0 ? t : 0.0) * (R[1] < 0 ? -1.0 : 1.0);
        t = (R[8] + 1.0) * 0.5;
        rz = sqrt(t > 0.0 ? t : 0.0) * (R[2] < 0 ? -1.0 : 1.0);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        const double nn = theta / sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * nn; r[1] = ry * nn; r[2] = rz * nn;
    } else {
        const double vth = theta / (2.0 * s);
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

// rotate_pose: global orientation <- Rodrigues^-1( Rz(-rot degrees) . R(global orientation) )
__device__ __forceinline__ void rotate_global(float* go, float rot_deg) {
    float R[9];
    tgm_aa_to_rotmat(go, R);
    const float ang = -3.14159265358979323846f * rot_deg / 180.f;       // torch: -np.pi * rot / 180. on a float32 tensor
    const float cs = cosf(ang), sn = sinf(ang);
    double M[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        M[c] = (double)(cs * R[c] + (-sn) * R[3 + c] + 0.f * R[6 + c]);
        M[3 + c] = (double)(sn * R[c] + cs * R[3 + c] + 0.f * R[6 + c]);
        M[6 + c] = (double)(0.f * R[c] + 0.f * R[3 + c] + 1.f * R[6 + c]);
    }
    double r[3];
    cv_rotmat_to_aa(M, r);
    go[0] = (float)r[0]; go[1] = (float)r[1]; go[2] = (float)r[2];
}

struct FlipPerm { uint8_t p[72]; };      // constants.SMPL_POSE_FLIP_PERM

// __getitem__: rows of the [N][82] store -> rotate -> flip
// Rows whose index falls outside [0, store_rows) are not touched in the store: the get returns NaNs for them, the set skips
// them, and both raise bit 0 of *status (device memory, nullable) - the bounds check needs no host round trip.
__global__ void __launch_bounds__(96) fits_get_kernel(const float* __restrict__ store, long long store_rows,
                                                      const long long* __restrict__ index,
                                                      const float* __restrict__ rot, const uint8_t* __restrict__ flipped,
                                                      const __grid_constant__ FlipPerm perm, float* __restrict__ pose,
                                                      float* __restrict__ betas, int* __restrict__ status, int batch) {
    __shared__ float row[82];
    const int b = blockIdx.x, t = threadIdx.x;
    const long long idx = index[b];
    if (idx < 0 || idx >= store_rows) {                          // block-uniform
        if (t == 0 && status) atomicOr(status, 1);
        if (t < 72) pose[(size_t)b * 72 + t] = __int_as_float(0x7fc00000);
        else if (t < 82) betas[(size_t)b * 10 + t - 72] = __int_as_float(0x7fc00000);
        return;
    }
    if (t < 82) row[t] = store[(size_t)idx * 82 + t];
    __syncthreads();
    if (t == 0) rotate_global(row, rot[b]);
    __syncthreads();
    if (t < 72) {
        float v = row[t];
        if (flipped[b]) {
            v = row[perm.p[t]];
            if (t % 3 != 0) v = -v;
        }
        pose[(size_t)b * 72 + t] = v;
    } else if (t < 82) {
        betas[(size_t)b * 10 + t - 72] = row[t];
    }
}

// __setitem__: flip -> rotate by -rot -> masked scatter into the store
__global__ void __launch_bounds__(96) fits_set_kernel(float* __restrict__ store, long long store_rows,
                                                      const long long* __restrict__ index,
                                                      const float* __restrict__ rot, const uint8_t* __restrict__ flipped,
                                                      const uint8_t* __restrict__ update, const __grid_constant__ FlipPerm perm,
                                                      const float* __restrict__ pose, const float* __restrict__ betas,
                                                      int* __restrict__ status, int batch) {
    __shared__ float row[82];
    const int b = blockIdx.x, t = threadIdx.x;
    const long long idx = index[b];
    if (idx < 0 || idx >= store_rows) {                          // reported whether or not the row was to be updated
        if (t == 0 && status) atomicOr(status, 1);
        return;
    }
    if (!update[b]) return;
    if (t < 72) {
        float v = pose[(size_t)b * 72 + t];
        if (flipped[b]) {
            v = pose[(size_t)b * 72 + perm.p[t]];
            if (t % 3 != 0) v = -v;
        }
        row[t] = v;
    } else if (t < 82) {
        row[t] = betas[(size_t)b * 10 + t - 72];
    }
    __syncthreads();
    if (t == 0) rotate_global(row, -rot[b]);
    __syncthreads();
    if (t < 82) store[(size_t)idx * 82 + t] = row[t];
}

// trainer.py:716-727: update = new_loss < old_loss (new_loss = mean over the 49 joints of the new reprojection loss);
// where it holds, the best-so-far loss / pose / betas / camera / joints rows are overwritten.
__global__ void __launch_bounds__(256) keep_better_kernel(const float* __restrict__ new_reproj, const float* __restrict__ new_pose,
                                                          const float* __restrict__ new_betas, const float* __restrict__ new_cam,
                                                          const float* __restrict__ new_joints, float* __restrict__ loss,
                                                          float* __restrict__ pose, float* __restrict__ betas, float* __restrict__ cam,
                                                          float* __restrict__ joints, uint8_t* __restrict__ update, int batch) {
    __shared__ int s_up;
    const int b = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        // mean over the 49 joints as an fp32 sum in index order / 49 (torch's reduction tree may differ in the last ulp,
        // which changes the decision only for ties within one ulp)
        float a = 0.f;
        for (int j = 0; j < 49; ++j) a += new_reproj[(size_t)b * 49 + j];
        const float m = a / 49.f;
        const int up = m < loss[b];
        s_up = up;
        update[b] = (uint8_t)up;
        if (up) loss[b] = m;
    }
    __syncthreads();
    if (!s_up) return;
    if (t < 72) pose[(size_t)b * 72 + t] = new_pose[(size_t)b * 72 + t];
    else if (t < 82) betas[(size_t)b * 10 + t - 72] = new_betas[(size_t)b * 10 + t - 72];
    else if (t < 85) cam[(size_t)b * 3 + t - 82] = new_cam[(size_t)b * 3 + t - 82];
    if (joints && new_joints && t < 147) joints[(size_t)b * 147 + t] = new_joints[(size_t)b * 147 + t];
}

// ---------------------------------------------------------------------------------------------------------------------
// train/trainer.py:187-199 (also :603-615, models/hmr.py:1708-1710, eval.py:245-247): weak-perspective camera [s, tx, ty]
// -> translation [tx, ty, 2 f / (img_res s + 1e-9)], perspective projection of the joints with identity rotation and zero
// camera centre, normalisation to [-1, 1] by img_res / 2.  One block per sample; same fp32 operation order as the eager ops.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) weak_persp_fwd_kernel(const float* __restrict__ joints, const float* __restrict__ cam, float focal,
                                                            float img_res, float* __restrict__ cam_t, float* __restrict__ kp, int npts) {
    const int b = blockIdx.x;
    const float s = cam[3 * b], tx = cam[3 * b + 1], ty = cam[3 * b + 2];
    // eager torch: two roundings in img_res * s + 1e-9 (no FMA), and scalar / tensor = tensor.reciprocal() * scalar
    const float tz = __fmul_rn(__frcp_rn(__fadd_rn(__fmul_rn(img_res, s), 1e-9f)), 2.f * focal);
    if (threadIdx.x == 0) { cam_t[3 * b] = tx; cam_t[3 * b + 1] = ty; cam_t[3 * b + 2] = tz; }
    const float half = img_res / 2.f;
    for (int n = threadIdx.x; n < npts; n += blockDim.x) {
        const float* J = joints + ((size_t)b * npts + n) * 3;
        const float x = J[0] + tx, y = J[1] + ty, z = J[2] + tz;
        kp[((size_t)b * npts + n) * 2 + 0] = (focal * (x / z)) / half;
        kp[((size_t)b * npts + n) * 2 + 1] = (focal * (y / z)) / half;
    }
}

// gradients w.r.t. the joints and the weak-perspective camera; g_cam_t (may be null) is an extra gradient on the translation
__global__ void __launch_bounds__(64) weak_persp_bwd_kernel(const float* __restrict__ joints, const float* __restrict__ cam, float focal,
                                                            float img_res, const float* __restrict__ g_kp, const float* __restrict__ g_cam_t,
                                                            float* __restrict__ g_joints, float* __restrict__ g_cam, int npts) {
    __shared__ float red[3][64];
    const int b = blockIdx.x, t = threadIdx.x;
    const float s = cam[3 * b], tx = cam[3 * b + 1], ty = cam[3 * b + 2];
    const float den = __fadd_rn(__fmul_rn(img_res, s), 1e-9f);
    const float tz = __fmul_rn(__frcp_rn(den), 2.f * focal);
    const float c = focal / (img_res / 2.f);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int n = t; n < npts; n += blockDim.x) {
        const float* J = joints + ((size_t)b * npts + n) * 3;
        const float x = J[0] + tx, y = J[1] + ty, z = J[2] + tz;
        const float gu = g_kp[((size_t)b * npts + n) * 2 + 0], gv = g_kp[((size_t)b * npts + n) * 2 + 1];
        const float gx = gu * c / z, gy = gv * c / z, gz = -(gu * c * x + gv * c * y) / (z * z);
        float* G = g_joints + ((size_t)b * npts + n) * 3;
        G[0] = gx; G[1] = gy; G[2] = gz;
        a0 += gx; a1 += gy; a2 += gz;
    }
    red[0][t] = a0; red[1][t] = a1; red[2][t] = a2;
    __syncthreads();
    if (t < 3) {
        float a = 0.f;
        for (int i = 0; i < 64; ++i) a += red[t][i];           // fixed order
        if (g_cam_t) a += g_cam_t[3 * b + t];
        if (t == 2) g_cam[3 * b + 0] = a * (-(2.f * focal) * img_res / (den * den));     // d tz / d s
        else g_cam[3 * b + 1 + t] = a;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
cudaError_t launch_rot6d_to_rotmat(const float* x, float* R, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    rot6d_to_rotmat_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, R, n);
    return cudaGetLastError();
}
cudaError_t launch_rotmat_to_aa(const float* R, float* aa, int n, int scrub_nan, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    rotmat_to_aa_kernel<<<(n + 255) / 256, 256, 0, st>>>(R, aa, n, scrub_nan);
    return cudaGetLastError();
}
cudaError_t launch_estimate_translation(const float* S, const float* kp, float focal, float img_size, float* trans, int batch,
                                        cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    estimate_translation_kernel<<<(batch + 127) / 128, 128, 0, st>>>(S, kp, focal, img_size, trans, batch);
    return cudaGetLastError();
}
static FlipPerm make_perm(const int* perm72) {
    FlipPerm p;
    for (int i = 0; i < 72; ++i) p.p[i] = (uint8_t)perm72[i];
    return p;
}
cudaError_t launch_fits_get(const float* store, long long store_rows, const long long* index, const float* rot, const uint8_t* flipped,
                            const int* perm72, float* pose, float* betas, int* status, int batch, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    fits_get_kernel<<<batch, 96, 0, st>>>(store, store_rows, index, rot, flipped, make_perm(perm72), pose, betas, status, batch);
    return cudaGetLastError();
}
cudaError_t launch_fits_set(float* store, long long store_rows, const long long* index, const float* rot, const uint8_t* flipped,
                            const uint8_t* update, const int* perm72, const float* pose, const float* betas, int* status, int batch,
                            cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    fits_set_kernel<<<batch, 96, 0, st>>>(store, store_rows, index, rot, flipped, update, make_perm(perm72), pose, betas, status, batch);
    return cudaGetLastError();
}
cudaError_t launch_keep_better(const float* new_reproj, const float* new_pose, const float* new_betas, const float* new_cam,
                               const float* new_joints, float* loss, float* pose, float* betas, float* cam, float* joints,
                               uint8_t* update, int batch, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    keep_better_kernel<<<batch, 256, 0, st>>>(new_reproj, new_pose, new_betas, new_cam, new_joints, loss, pose, betas, cam, joints,
                                              update, batch);
    return cudaGetLastError();
}

cudaError_t launch_weak_persp_fwd(const float* joints, const float* cam, float focal, float img_res, float* cam_t, float* kp, int batch,
                                  int npts, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    weak_persp_fwd_kernel<<<batch, 64, 0, st>>>(joints, cam, focal, img_res, cam_t, kp, npts);
    return cudaGetLastError();
}
cudaError_t launch_weak_persp_bwd(const float* joints, const float* cam, float focal, float img_res, const float* g_kp,
                                  const float* g_cam_t, float* g_joints, float* g_cam, int batch, int npts, cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    weak_persp_bwd_kernel<<<batch, 64, 0, st>>>(joints, cam, focal, img_res, g_kp, g_cam_t, g_joints, g_cam, npts);
    return cudaGetLastError();
}

}  // namespace smplb200
