// Host interface of the tcgen05 LBS vertex kernel (lbs_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "smpl_common.h"

namespace smplb200 {

struct alignas(64) TcMaps {       // TMA tensor maps, passed to the kernel as a __grid_constant__ parameter
    CUtensorMap x_hi, x_lo;       // [B][224]       blend coefficients, hi / lo tf32 parts
    CUtensorMap b_hi, b_lo;       // [20736][224]   basis^T (row = vertex coordinate)
    CUtensorMap ae_hi, ae_lo;     // [12][B][32]    skinning transforms, entry-major, 24 joints padded to 32
    CUtensorMap w_hi, w_lo;       // [6912][32]     skinning weights, 24 joints padded to 32
};

struct TcOperands {               // per-call device buffers written by the pose kernels
    float* x_hi;
    float* x_lo;
    float* ae_hi;
    float* ae_lo;
};

constexpr int kTcVertRowsPad = 6912;       // 216 vertex tiles x 32

bool tc_make_constant_maps(TcMaps* maps, const float* basisT_hi, const float* basisT_lo, const float* w_hi, const float* w_lo);
cudaError_t launch_vertex_forward_tc(const TcMaps& constant_maps, const TcOperands& op, float* verts, float* vposed, int batch,
                                     cudaStream_t stream);

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
