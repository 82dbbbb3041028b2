// Shared definitions of the B200 SMPLify / SMPL library: problem sizes, the device-side
// view of the immutable model constants, and the host/device portability macros that let
// the per-sample "tile" code be compiled for the host by the TEST-ONLY emulation build
// (tests/emu, -DSMPLB200_EMU).  The product library never runs that host path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
// Phases that are compiled as real functions: inside the one big fit kernel the register allocator otherwise pads the
// streamed-constant GEMM loops with register moves (64 MOVs per 192 FFMA2 against 4 when the loop is compiled alone).
#define SB_HD_CALL __host__ __device__ __noinline__
#else
#define SB_HD inline
#define SB_HD_CALL inline
#endif

namespace smplb200 {

constexpr int kJoints = 24;          // SMPL kinematic joints
constexpr int kBetas = 10;
constexpr int kPoseFeat = 207;       // 9 * 23
constexpr int kX = 218;              // x = [1, betas(10), pose_feature(207)]
constexpr int kXPad = 224;           // padded leading dimension of x
constexpr int kVerts = 6890;
constexpr int kCols = 20670;         // 3 * 6890 vertex coordinates
constexpr int kColsPad = 20736;      // basis row stride (multiple of 64 floats)
constexpr int kExtra = 9;            // J_regressor_extra joints (54-joint indices 45..53)
constexpr int kPicks = 11;           // selected-vertex joints actually referenced by joint_map (24..34)
constexpr int kSelVerts = 21;        // all VertexJointSelector vertices (24..44)
constexpr int kQ = 681;              // 9*24*3 folded extra-joint terms + 11*3 picked-vertex coordinates
constexpr int kQPad = 704;
constexpr int kQPickBase = 648;
constexpr int kOut = 49;             // joints after joint_map
constexpr int kSrc = 54;             // 24 chain + 21 selected vertices + 9 extra
constexpr int kGauss = 8;
constexpr int kPriorDim = 69;
constexpr int kPriorPad = 72;         // padded row/column count of the prior matrices
constexpr int kParams = 82;          // 72 pose + 10 betas
constexpr int kMaxLevels = 24;

// Device-visible (or, in the emulation build, host-visible) model constants.
struct ModelView {
    const float* basis;      // [kXPad][kColsPad]  row m of x -> vertex coordinates (3v+c); row 0 = v_template
    const float* basisT;     // [kColsPad][kXPad]  transpose of basis (backward blend GEMM)
    const float* weights;    // [kVerts][24]       skinning weights
    const float* Cf;         // [kXPad][kQPad]     folded joint basis, m-major   (forward)
    const float* CfT;        // [kQPad][kXPad]     same, n-major                 (backward)
    const float* wkj;        // [9][24]            sum_v Jextra[k,v] W[v,j]
    const float* Wp;         // [11][24]           skinning weights of the picked vertices
    const float* J0;         // [24][3]            J_regressor . v_template
    const float* JS;         // [24][3][10]        J_regressor . shapedirs
    const float* gmm_means;  // [8][72]            zero padded
    const float* gmm_prec;   // [8][72 j][72 i]    symmetrised precision, i fastest, zero padded
    const float* gmm_pmean;  // [8][72]            prec_sym . mean
    const float* gmm_lognll; // [8]                log(nll_weights)
    const float* pg_prior;   // packed A operands of the pair kernel's GEMMs (kPg*Floats): symmetrised precisions,
    const float* pg_fwd;     //   Cf^T (rows n, k = m),
    const float* pg_bwd;     //   Cf   (rows m, k = n)
    int32_t pick_vid[kSelVerts];
    int8_t parents[kJoints];
    uint8_t joint_map[kOut];
    int8_t inv_map[kSrc][2];           // outputs fed by each source joint (-1 = none)
    uint8_t level_order[kJoints];      // joints sorted by tree depth
    uint8_t level_start[kMaxLevels + 1];
    uint8_t child_start[kJoints + 1];
    uint8_t child_list[kJoints];
    int32_t num_levels;
    float sigma_src[kSrc];             // sum of the affine weights behind each source joint (1 for chain joints)
    int32_t angle_ids[4];              // body_pose entries of the angle prior
    float angle_signs[4];
    uint8_t ign_joints[8];             // output joints whose confidence the body stage zeroes
    int32_t num_ign;
    uint8_t cam_op[4], cam_gt[4];      // output joints of the camera stage (OpenPose / ground truth)
};

struct AdamScalars {           // per-iteration host-computed scalars (torch computes them in float64)
    float step_size;           // lr / (1 - beta1^t)
    float bc2_sqrt;            // sqrt(1 - beta2^t)
};

// Shapes of the pair fit kernel's tensor-core GEMMs (pair_gemm.cuh): 256-row tiles (128 rows per CTA of the pair), chunks
// of 32 k values.  Packed constant arrays: [tile][half][k chunk][8 float4 of a row][128 rows] float4.
constexpr int kPgChunk = 32;
constexpr int kPgPriorTiles = 3, kPgPriorChunks = 3;                       // 576 rows (g, i), K = 72 (zero padded to 96)
constexpr int kPgFwdTiles = 3, kPgFwdChunks = kXPad / kPgChunk;            // 704 rows n, K = 224
constexpr int kPgBwdTiles = 1, kPgBwdChunks = kQPad / kPgChunk;            // 224 rows m, K = 704
constexpr size_t kPgPriorFloats = (size_t)kPgPriorTiles * 2 * kPgPriorChunks * kPgChunk * 128;
constexpr size_t kPgFwdFloats = (size_t)kPgFwdTiles * 2 * kPgFwdChunks * kPgChunk * 128;
constexpr size_t kPgBwdFloats = (size_t)kPgBwdTiles * 2 * kPgBwdChunks * kPgChunk * 128;

constexpr int kFitTileThreads = 384; // threads of the standard tile (fit and pose kernels)
constexpr int kMaxIters = 256;     // per stage; bounds the in-kernel Adam scalar table

}  // namespace smplb200
