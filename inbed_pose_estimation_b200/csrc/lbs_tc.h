// Host interface of the tcgen05 LBS vertex kernels (lbs_tc.cu): 3xTF32 GEMMs fed by TMA with fp32 accumulators in
// TMEM, for the forward and the backward pass of the per-vertex half of SMPL (smplx lbs rows a6, a9, a11).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "smpl_common.h"

namespace smplb200 {

constexpr int kTcVertRowsPad = 6912;        // 54 vertex blocks x 128 = 216 k-blocks x 32
constexpr int kVpPitch = kColsPad;          // row pitch (floats) of the v_posed / dvp buffers: 20736, 16-byte aligned rows
constexpr int kAeRow = 12 * 32;             // floats per sample of the skinning-transform operand: [12 entries][24 joints + 8 pad]
constexpr int kMaxSplitA = 32;              // vertex-range splits of the dA kernel (bounds the tensor-core accumulation chains)
constexpr int kMaxSplitX = 16;              // K splits of the dx GEMM

struct alignas(64) TcConstMaps {  // TMA tensor maps of the immutable model operands (hi / lo tf32 parts)
    CUtensorMap bT_hi, bT_lo;     // basis^T [20736][224]  box 256 x 32   B operand of the blend-shape GEMM
    CUtensorMap bm_hi, bm_lo;     // basis   [224][20736]  box 224 x 32   B operand of the dx GEMM
    CUtensorMap bm_hi_half, bm_lo_half;   // same tensors, box 112 x 32: the half tile one CTA of a cluster pair multicasts
    CUtensorMap w_hi, w_lo;       // W       [6912][32]    box 128 x 32   A operand of the skinning GEMM
    CUtensorMap wT_hi, wT_lo;     // W^T     [32][6912]    box  32 x 32   B operand of the dA GEMM
};

struct TcOperands {               // per-call device buffers written by the pose kernels
    float* x_hi;                  // [B][224]      blend coefficients x = [1, beta, pose_feature], tf32 hi part
    float* x_lo;                  //               x - hi
    float* ae_hi;                 // [B][12][32]   skinning transforms, entry-major per sample, 24 joints padded to 32
    float* ae_lo;
};

bool tc_make_constant_maps(TcConstMaps* maps, const float* basisT_hi, const float* basisT_lo, const float* basis_hi,
                           const float* basis_lo, const float* w_hi, const float* w_lo, const float* wT_hi, const float* wT_lo);

// v_posed[B][20736] = x . basis            (tcgen05, M = samples, N = vertex coordinates, K = 224)
cudaError_t launch_blend_gemm(const TcConstMaps& cm, const float* x_hi, const float* x_lo, float* vposed, int batch, cudaStream_t stream);
// verts[b][v] = (W . A_b)[v] [v_posed[b][v]; 1]      (tcgen05 skinning GEMM, M = vertices, N = (sample, entry), K = 24)
cudaError_t launch_skin_forward(const TcConstMaps& cm, const float* ae_hi, const float* ae_lo, const float* vposed, float* verts,
                                int batch, cudaStream_t stream);
// dvp[b][v] = (W . A_b)[v]^R^T dverts[b][v], written in fp32 [B][20736]; also copies dverts into the padded, 16-byte aligned
// layout [B][20736] the dA kernel's TMA loads need (the caller's [B][6890][3] rows are only 8-byte aligned)
cudaError_t launch_skin_backward(const TcConstMaps& cm, const float* ae_hi, const float* ae_lo, const float* dverts, float* dvp,
                                 float* dverts_padded, int batch, cudaStream_t stream);
// dx_part[split][B][224] = dvp . basis^T over the split's K range   (tcgen05, K = 20736 split nsplit ways; dvp is split
// into its tf32 hi / lo parts on chip)
cudaError_t launch_dx_gemm(const TcConstMaps& cm, const float* dvp, float* dx_part, int batch, int nsplit, cudaStream_t stream);
// dA_part[split][B][12][24] = sum_v W[v][j] dverts[b][v][r] [v_posed[b][v]; 1][c]   (tcgen05, operands generated on chip)
cudaError_t launch_dA(const TcConstMaps& cm, const float* dverts_padded, const float* vposed, float* dA_part, int batch, int nsplit,
                      cudaStream_t stream);
int tc_dA_splits(int batch);
int tc_dx_splits(int batch);

// hi = x rounded to tf32 (10 explicit mantissa bits, round to nearest even); lo = x - hi is exact in fp32
SB_HD float tf32_round(float x) {
#if defined(__CUDA_ARCH__)
    unsigned u = __float_as_uint(x);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;
    return __uint_as_float(u);
#else
    union { float f; unsigned u; } c;
    c.f = x;
    c.u = (c.u + 0xFFFu + ((c.u >> 13) & 1u)) & 0xFFFFE000u;
    return c.f;
#endif
}

}  // namespace smplb200
