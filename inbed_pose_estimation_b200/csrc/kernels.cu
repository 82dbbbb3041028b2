// CUDA kernels (sm_100a) of the SMPLify / SMPL hot path and their host launchers.
//
//   smplify_fit_kernel<S>        one CTA = S samples, the whole two-stage fit in one launch
//   pose_forward_kernel<S>       per-sample half of SMPL.forward (joints, skinning transforms, blend coefficients)
//   pose_backward_kernel<S>      its gradient
//   (the per-vertex half - blend shapes, skinning and their gradients - is the tcgen05 path in lbs_tc.cu)
//   quat_rodrigues_{fwd,bwd}     utils/geometry.py:9-45
//   projection_{fwd,bwd}         utils/geometry.py:79-107
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "fit_driver.cuh"
#include "fit_pair.cuh"
#include "fit_split.cuh"
#include "launch.h"

namespace smplb200 {

// ------------------------------------------------------------------------------------------------
// tile kernels
// ------------------------------------------------------------------------------------------------
template <int S, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) smplify_fit_kernel(const __grid_constant__ ModelView M,
                                                               const __grid_constant__ FitParams P) {
    extern __shared__ __align__(16) float sm[];
    fit_tile<S>(M, P, blockIdx.x * S, sm);
}

// Two tile sizes in one launch: CTAs [0, n_a) fit SA samples each, the rest SB samples each, starting where the first group
// ends.  The hardware hands out CTAs in index order, so the larger tiles form the first wave(s) and the smaller ones fill
// the last wave: a batch of 4096 runs as 148 x 16 + 144 x 12 samples - two full waves - instead of 256 x 16 = 1.73 waves.
template <int SA, int SB>
__global__ void __launch_bounds__(kFitThreads, 1) smplify_fit_mixed_kernel(const __grid_constant__ ModelView M,
                                                                           const __grid_constant__ FitParams P, int n_a) {
    extern __shared__ __align__(16) float sm[];
    const int blk = blockIdx.x;                                  // block-uniform branch
    if (blk < n_a) fit_tile<SA>(M, P, blk * SA, sm);
    else fit_tile<SB>(M, P, n_a * SA + (blk - n_a) * SB, sm);
}

// Small batches: a cluster of C CTAs per 4-sample tile, the three per-iteration GEMMs split by output rows (fit_split.cuh).
// The cluster size is a launch attribute (cudaLaunchKernelEx).
template <int C>
__global__ void __launch_bounds__(kFitThreads, 1) smplify_fit_split_kernel(const __grid_constant__ ModelView M,
                                                                           const __grid_constant__ FitParams P) {
    extern __shared__ __align__(16) float sm[];
    fit_split_tile<C>(M, P, (int)(blockIdx.x / C) * kSplitS, sm);
}

// Pairs of CTAs (2-CTA clusters) that share their GEMMs on the tensor cores (fit_pair.cuh): pairs [0, n_a) fit 2 x SA samples,
// the others 2 x SB, starting where the first group ends.  One CTA per SM; a wave is sms / 2 pairs.
template <int SA, int SB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pg::kThreads, 1)
    smplify_fit_pair_kernel(const __grid_constant__ ModelView M, const __grid_constant__ FitParams P, int n_a) {
    extern __shared__ __align__(16) float sm_raw[];
    float* sm = sm_raw + ((1024u - (tc::smem_u32(sm_raw) & 1023u)) & 1023u) / 4;          // SWIZZLE_128B atoms: 1024-byte aligned
    const uint32_t rank = tc::cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    if (pair < n_a) fit_pair_tile<SA>(M, P, pair * 2 * SA + (int)rank * SA, sm, rank);
    else fit_pair_tile<SB>(M, P, n_a * 2 * SA + (pair - n_a) * 2 * SB + (int)rank * SB, sm, rank);
}

template <int S>
__global__ void __launch_bounds__(kPoseThreads) pose_forward_kernel(const __grid_constant__ ModelView M,
                                                                    const __grid_constant__ PoseParams P) {
    extern __shared__ __align__(16) float sm[];
    pose_forward_tile<S>(M, P, blockIdx.x * S, sm);
}

template <int S>
__global__ void __launch_bounds__(kPoseThreads) prior_terms_kernel(const __grid_constant__ ModelView M,
                                                                   const __grid_constant__ PriorParams P) {
    extern __shared__ __align__(16) float sm[];
    prior_tile<S>(M, P, blockIdx.x * S, sm);
}

template <int S>
__global__ void __launch_bounds__(kPoseThreads) pose_backward_kernel(const __grid_constant__ ModelView M,
                                                                     const __grid_constant__ PoseParams P) {
    extern __shared__ __align__(16) float sm[];
    pose_backward_tile<S>(M, P, blockIdx.x * S, sm);
}

// SM count of the CURRENT device, cached per device index (a process may drive several GPUs).
int device_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;                                  // benign race: every thread writes the same value
    }
    return n;
}

template <int S>
static size_t tile_smem_bytes() { return (size_t)TileLayout<S>::SMEM_FLOATS * sizeof(float); }

template <typename K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#if defined(PG_TRACE)
cudaError_t debug_pg_trace(long long* out512) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out512, pg::g_pg_trace, sizeof(long long) * 512);
    return e;
}
#endif
#if defined(SMPLB200_PHASE_CLOCKS)
cudaError_t debug_phase_clocks(unsigned long long* out32, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && out32) e = cudaMemcpyFromSymbol(out32, g_phase_clocks, sizeof(unsigned long long) * 32);
    if (e == cudaSuccess && reset) {
        unsigned long long z[32] = {0};
        e = cudaMemcpyToSymbol(g_phase_clocks, z, sizeof(z));
    }
    return e;
}
#endif

template <int S, int NT, int MINB>
static cudaError_t launch_fit_variant(const ModelView& M, const FitParams& P, cudaStream_t stream) {
    cudaError_t e = opt_in_smem(smplify_fit_kernel<S, NT, MINB>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    smplify_fit_kernel<S, NT, MINB><<<(P.batch + S - 1) / S, NT, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}

// Tile plan of a batch: n16 tiles of 16 samples followed by n_small tiles of `small` samples (4, 8 or 12).  A wave is one
// tile per SM and lasts as long as its tile size dictates (B200, 100 + 100 iterations: 2.6 / 3.5 / 4.7 / 5.3 ms for
// 4 / 8 / 12 / 16 samples, profiles/fit_kernel_r1.md), so: whole waves of 16s first, then the remainder as ONE wave of the
// smallest tile size that still covers it (at most `small` samples per SM), else as one more wave of 16s.
void plan_fit_tiles(int batch, int sms, int* n16, int* small, int* n_small) {
    const int full = batch / (16 * sms);
    const int rest = batch - full * 16 * sms;                    // < 16 * sms
    *n16 = full * sms;
    *small = 0;
    *n_small = 0;
    if (rest == 0) return;
    for (int s = 4; s <= 12; s += 4)
        if ((rest + s - 1) / s <= sms) { *small = s; *n_small = (rest + s - 1) / s; return; }
    *n16 += (rest + 15) / 16;
}

template <int SB>
static cudaError_t launch_fit_mixed(const ModelView& M, const FitParams& P, int n16, int n_small, cudaStream_t stream) {
    const size_t smem = tile_smem_bytes<16>() > tile_smem_bytes<SB>() ? tile_smem_bytes<16>() : tile_smem_bytes<SB>();
    cudaError_t e = opt_in_smem(smplify_fit_mixed_kernel<16, SB>, smem);
    if (e != cudaSuccess) return e;
    smplify_fit_mixed_kernel<16, SB><<<n16 + n_small, kFitThreads, smem, stream>>>(M, P, n16);
    return cudaGetLastError();
}

// Pair plan of a batch: n16 pairs of 2 x 16 samples in whole waves of sms / 2 pairs, the remainder as pairs of 2 x 12 when
// one wave of those covers it, else as more pairs of 2 x 16.
void plan_fit_pairs(int batch, int sms, int* n16, int* n12) {
    const int wave = sms / 2;
    const int full = batch / (32 * wave);
    const int rest = batch - full * 32 * wave;
    *n16 = full * wave;
    *n12 = 0;
    if (rest == 0) return;
    if ((rest + 23) / 24 <= wave) *n12 = (rest + 23) / 24;
    else *n16 += (rest + 31) / 32;
}

static cudaError_t launch_fit_pairs(const ModelView& M, const FitParams& P, int sms, cudaStream_t stream) {
    int n16 = 0, n12 = 0;
    plan_fit_pairs(P.batch, sms, &n16, &n12);
    const size_t smem = (size_t)(PairLayout<16>::SMEM_FLOATS > PairLayout<12>::SMEM_FLOATS ? PairLayout<16>::SMEM_FLOATS
                                                                                          : PairLayout<12>::SMEM_FLOATS) * sizeof(float) + 1024;
    cudaError_t e = opt_in_smem(smplify_fit_pair_kernel<16, 12>, smem);
    if (e != cudaSuccess) return e;
    smplify_fit_pair_kernel<16, 12><<<2 * (n16 + n12), pg::kThreads, smem, stream>>>(M, P, n16);
    return cudaGetLastError();
}

// Cluster size of the split kernel for a batch: the largest of 8 / 4 / 2 CTAs per 4-sample tile whose clusters are all resident
// at once; 0 = the batch is too large for it (the 4- / 8- / 12-sample tiles take over).  `capacity[i]` = clusters of 8 / 4 / 2
// CTAs the device holds at a time.
static int plan_fit_split_caps(int batch, const int capacity[3]) {
    const int tiles = (batch + kSplitS - 1) / kSplitS;
    for (int i = 0, c = 8; c >= 2; c /= 2, ++i)
        if (tiles <= capacity[i]) return c;
    return 0;
}
// ... by SM count alone (an upper bound: sms / C clusters)
int plan_fit_split(int batch, int sms) {
    const int cap[3] = {sms / 8, sms / 4, sms / 2};
    return plan_fit_split_caps(batch, cap);
}

template <int C>
static void fill_split_launch(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int clusters, cudaStream_t stream) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)(clusters * C));
    cfg.blockDim = dim3(kFitThreads);
    cfg.dynamicSmemBytes = (size_t)SplitLayout::SMEM_FLOATS * sizeof(float);
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

// clusters of C CTAs of the split kernel the CURRENT device holds at a time (one CTA per SM; a cluster lives inside one GPC, so
// this is less than sms / C: 16 clusters of 8 would need 128 of the 148 SMs in whole groups of 8 per GPC - measured on B200: a
// batch of 64 planned by SM count alone ran as two waves, 3.67 instead of 1.88 ms)
template <int C>
static int split_cluster_capacity() {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    fill_split_launch<C>(cfg, attr, 1, nullptr);
    int n = 0;
    if (opt_in_smem(smplify_fit_split_kernel<C>, cfg.dynamicSmemBytes) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveClusters(&n, smplify_fit_split_kernel<C>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
// ... for the current device, cached per device index
int plan_fit_split_device(int batch) {
    static int cache[64][3];
    static bool known[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    if (!known[dev]) {
        cache[dev][0] = split_cluster_capacity<8>();
        cache[dev][1] = split_cluster_capacity<4>();
        cache[dev][2] = split_cluster_capacity<2>();
        known[dev] = true;                               // benign race: every thread writes the same values
    }
    return plan_fit_split_caps(batch, cache[dev]);
}

template <int C>
static cudaError_t launch_fit_split(const ModelView& M, const FitParams& P, cudaStream_t stream) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    fill_split_launch<C>(cfg, attr, (P.batch + kSplitS - 1) / kSplitS, stream);
    cudaError_t e = opt_in_smem(smplify_fit_split_kernel<C>, cfg.dynamicSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, smplify_fit_split_kernel<C>, M, P);
}

// Batches from this size on run on the pair kernel (tensor-core GEMMs shared by two CTAs); smaller ones on the 4- / 8-sample
// tiles of the CUDA-core kernel, which fill more SMs.
constexpr int kPairMinBatch = 1024;

static int fit_variant() {
    static const int variant = [] { const char* v = getenv("SMPLB200_FIT_VARIANT"); return v ? atoi(v) : 0; }();   // experiments only
    return variant;
}
int fit_uses_pairs(int batch, int num_iters) {
    const int variant = fit_variant();
    return (variant == 10 || (variant == 0 && num_iters > 0 && batch >= kPairMinBatch)) ? 1 : 0;
}

cudaError_t launch_fit(const ModelView& M, const FitParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    const int variant = fit_variant();
    if (fit_uses_pairs(P.batch, P.num_iters) && M.pg_fwd != nullptr) return launch_fit_pairs(M, P, device_sm_count(), stream);
    if ((variant == 0 || variant == 12) && P.num_iters > 0) {
        const int c = plan_fit_split_device(P.batch);
        if (c == 8) return launch_fit_split<8>(M, P, stream);
        if (c == 4) return launch_fit_split<4>(M, P, stream);
        if (c == 2) return launch_fit_split<2>(M, P, stream);
    }
    if (variant == 1) return launch_fit_variant<8, 192, 2>(M, P, stream);
    if (variant == 2) return launch_fit_variant<8, 256, 2>(M, P, stream);
    if (variant == 3) return launch_fit_variant<16, 384, 1>(M, P, stream);
    if (variant == 4) return launch_fit_variant<12, 384, 1>(M, P, stream);
    if (variant == 6) return launch_fit_variant<8, 384, 1>(M, P, stream);
    if (variant == 7) return launch_fit_variant<4, 384, 1>(M, P, stream);
    const int sms = device_sm_count();
    int n16 = 0, small = 0, n_small = 0;
    plan_fit_tiles(P.batch, sms, &n16, &small, &n_small);
    if (n16 == 0) {
        if (small == 4) return launch_fit_variant<4, kFitThreads, 1>(M, P, stream);
        if (small == 8) return launch_fit_variant<8, kFitThreads, 1>(M, P, stream);
        return launch_fit_variant<12, kFitThreads, 1>(M, P, stream);
    }
    if (n_small == 0) return launch_fit_variant<16, kFitThreads, 1>(M, P, stream);
    if (small == 4) return launch_fit_mixed<4>(M, P, n16, n_small, stream);
    if (small == 8) return launch_fit_mixed<8>(M, P, n16, n_small, stream);
    return launch_fit_mixed<12>(M, P, n16, n_small, stream);
}

template <int S>
static cudaError_t launch_pose_forward_s(const ModelView& M, const PoseParams& P, cudaStream_t stream) {
    cudaError_t e = opt_in_smem(pose_forward_kernel<S>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    pose_forward_kernel<S><<<(P.batch + S - 1) / S, kPoseThreads, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}
template <int S>
static cudaError_t launch_pose_backward_s(const ModelView& M, const PoseParams& P, cudaStream_t stream) {
    cudaError_t e = opt_in_smem(pose_backward_kernel<S>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    pose_backward_kernel<S><<<(P.batch + S - 1) / S, kPoseThreads, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}

template <int S>
static cudaError_t launch_prior_terms_s(const ModelView& M, const PriorParams& P, cudaStream_t stream) {
    cudaError_t e = opt_in_smem(prior_terms_kernel<S>, tile_smem_bytes<S>());
    if (e != cudaSuccess) return e;
    prior_terms_kernel<S><<<(P.batch + S - 1) / S, kPoseThreads, tile_smem_bytes<S>(), stream>>>(M, P);
    return cudaGetLastError();
}
// Both GEMM forms of the prior phase the fit kernel uses are reachable: the K-split form of the one-group tiles (S = 8) and
// the two-sample-group form of the 16-sample tiles.
cudaError_t launch_prior_terms(const ModelView& M, const PriorParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    return P.batch >= 16 * 128 ? launch_prior_terms_s<16>(M, P, stream) : launch_prior_terms_s<8>(M, P, stream);
}

// 16 samples per CTA halve the folded-basis bytes streamed from L2 per sample; small batches use 8 to fill more SMs.
cudaError_t launch_pose_forward(const ModelView& M, const PoseParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    return P.batch >= 16 * 128 ? launch_pose_forward_s<16>(M, P, stream) : launch_pose_forward_s<8>(M, P, stream);
}

cudaError_t launch_pose_backward(const ModelView& M, const PoseParams& P, cudaStream_t stream) {
    if (P.batch <= 0) return cudaSuccess;
    return P.batch >= 16 * 128 ? launch_pose_backward_s<16>(M, P, stream) : launch_pose_backward_s<8>(M, P, stream);
}

// ------------------------------------------------------------------------------------------------
// utils/geometry.py:9-45   batch_rodrigues (quaternion form) forward / backward
// ------------------------------------------------------------------------------------------------
__global__ void quat_rodrigues_fwd_kernel(const float* __restrict__ theta, float* __restrict__ rot, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rx = theta[3 * i], ry = theta[3 * i + 1], rz = theta[3 * i + 2];
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
    float* R = rot + 9 * (size_t)i;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

__global__ void quat_rodrigues_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ grot,
                                          float* __restrict__ gtheta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float rx = theta[3 * i], ry = theta[3 * i + 1], rz = theta[3 * i + 2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float L = sqrtf(ax * ax + ay * ay + az * az);
    const float nx = rx / L, ny = ry / L, nz = rz / L;
    float sh, ch;
    sincosf(L * 0.5f, &sh, &ch);
    const float qw = ch, qx = sh * nx, qy = sh * ny, qz = sh * nz;
    const float N = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
    const float w = qw / N, x = qx / N, y = qy / N, z = qz / N;
    const float* g = grot + 9 * (size_t)i;
    const float dw = 2 * w * (g[0] + g[4] + g[8]) + 2 * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
    const float dx = 2 * x * (g[0] - g[4] - g[8]) + 2 * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
    const float dy = 2 * y * (-g[0] + g[4] - g[8]) + 2 * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
    const float dz = 2 * z * (-g[0] - g[4] + g[8]) + 2 * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
    const float dot = w * dw + x * dx + y * dy + z * dz;        // through q / |q|
    const float dqw = (dw - w * dot) / N, dqx = (dx - x * dot) / N, dqy = (dy - y * dot) / N, dqz = (dz - z * dot) / N;
    const float dh = -sh * dqw + ch * (nx * dqx + ny * dqy + nz * dqz);
    const float dnx = sh * dqx, dny = sh * dqy, dnz = sh * dqz;
    float dL = 0.5f * dh;
    dL -= (dnx * rx + dny * ry + dnz * rz) / (L * L);
    gtheta[3 * i + 0] = dnx / L + dL * ax / L;
    gtheta[3 * i + 1] = dny / L + dL * ay / L;
    gtheta[3 * i + 2] = dnz / L + dL * az / L;
}

cudaError_t launch_quat_rodrigues_fwd(const float* theta, float* rot, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    quat_rodrigues_fwd_kernel<<<(n + 255) / 256, 256, 0, st>>>(theta, rot, n);
    return cudaGetLastError();
}
cudaError_t launch_quat_rodrigues_bwd(const float* theta, const float* grot, float* gtheta, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    quat_rodrigues_bwd_kernel<<<(n + 255) / 256, 256, 0, st>>>(theta, grot, gtheta, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// utils/geometry.py:79-114  perspective_projection forward / backward.  out_3d (:108-114; callers train/trainer.py:621-626,
// models/hmr.py:1720, eval.py:255): a third channel holds the camera-space depth (row 2 of K applied to the un-normalised
// point), the first two are unchanged.
// ------------------------------------------------------------------------------------------------
__global__ void projection_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ rot, const float* __restrict__ tr,
                                      const float* __restrict__ focal, int focal_per_batch, const float* __restrict__ cen,
                                      float* __restrict__ out, int out_3d, int batch, int npts) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)batch * npts) return;
    const int b = (int)(i / npts);
    const float* R = rot + 9 * (size_t)b;
    const float X = pts[3 * i], Y = pts[3 * i + 1], Z = pts[3 * i + 2];
    const float Px = (R[0] * X + R[1] * Y + R[2] * Z) + tr[3 * b + 0];
    const float Py = (R[3] * X + R[4] * Y + R[5] * Z) + tr[3 * b + 1];
    const float Pz = (R[6] * X + R[7] * Y + R[8] * Z) + tr[3 * b + 2];
    const float f = focal[focal_per_batch ? b : 0];
    const size_t o = (out_3d ? 3 : 2) * i;
    out[o + 0] = f * (Px / Pz) + cen[2 * b + 0];
    out[o + 1] = f * (Py / Pz) + cen[2 * b + 1];
    if (out_3d) out[o + 2] = Pz;
}

// one CTA per batch element; d_rot / d_tr are reduced over the points in shared memory
__global__ void __launch_bounds__(128) projection_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ rot,
                                                             const float* __restrict__ tr, const float* __restrict__ focal,
                                                             int focal_per_batch, const float* __restrict__ gout, int out_3d,
                                                             float* __restrict__ gpts, float* __restrict__ grot,
                                                             float* __restrict__ gtr, int npts) {
    __shared__ float red[4][12];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* R = rot + 9 * (size_t)b;
    const float f = focal[focal_per_batch ? b : 0];
    float acc[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) acc[e] = 0.f;
    for (int n = tid; n < npts; n += 128) {
        const size_t i = (size_t)b * npts + n;
        const float X = pts[3 * i], Y = pts[3 * i + 1], Z = pts[3 * i + 2];
        const float Px = (R[0] * X + R[1] * Y + R[2] * Z) + tr[3 * b + 0];
        const float Py = (R[3] * X + R[4] * Y + R[5] * Z) + tr[3 * b + 1];
        const float Pz = (R[6] * X + R[7] * Y + R[8] * Z) + tr[3 * b + 2];
        const size_t go = (out_3d ? 3 : 2) * i;
        const float gu = gout[go] * f, gv = gout[go + 1] * f;
        const float dP[3] = {gu / Pz, gv / Pz, -(gu * (Px / Pz) + gv * (Py / Pz)) / Pz + (out_3d ? gout[go + 2] : 0.f)};
        gpts[3 * i + 0] = R[0] * dP[0] + R[3] * dP[1] + R[6] * dP[2];
        gpts[3 * i + 1] = R[1] * dP[0] + R[4] * dP[1] + R[7] * dP[2];
        gpts[3 * i + 2] = R[2] * dP[0] + R[5] * dP[1] + R[8] * dP[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            acc[r * 3 + 0] += dP[r] * X; acc[r * 3 + 1] += dP[r] * Y; acc[r * 3 + 2] += dP[r] * Z;
            acc[9 + r] += dP[r];
        }
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        float v = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][e] = v;
    }
    __syncthreads();
    if (tid < 12) {
        const float v = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
        if (tid < 9) grot[9 * (size_t)b + tid] = v;
        else gtr[3 * (size_t)b + tid - 9] = v;
    }
}

cudaError_t launch_projection_fwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* cen, float* out, int out_3d, int batch, int npts, cudaStream_t st) {
    const size_t n = (size_t)batch * npts;
    if (n == 0) return cudaSuccess;
    projection_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pts, rot, tr, focal, focal_per_batch, cen, out, out_3d, batch,
                                                                       npts);
    return cudaGetLastError();
}
cudaError_t launch_projection_bwd(const float* pts, const float* rot, const float* tr, const float* focal, int focal_per_batch,
                                  const float* gout, int out_3d, float* gpts, float* grot, float* gtr, int batch, int npts,
                                  cudaStream_t st) {
    if (batch <= 0) return cudaSuccess;
    projection_bwd_kernel<<<batch, 128, 0, st>>>(pts, rot, tr, focal, focal_per_batch, gout, out_3d, gpts, grot, gtr, npts);
    return cudaGetLastError();
}

}  // namespace smplb200
