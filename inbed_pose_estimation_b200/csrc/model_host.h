// Host-side construction of the model constants (float64 folding of the joint regressors).
// Shared by the product library (uploads the arrays to the device) and by the TEST-ONLY
// emulation build (keeps them on the host).
#pragma once
#include <string>
#include <vector>
#include "smpl_common.h"
#include "../../include/smplify_b200.h"

namespace smplb200 {

struct HostModel {
    std::vector<float> basis, basisT, weights, Cf, CfT, wkj, Wp, J0, JS;
    std::vector<float> gmm_means, gmm_prec, gmm_pmean, gmm_lognll;
    std::vector<float> pg_prior, pg_fwd, pg_bwd;          // packed A operands of the pair kernel (smpl_common.h kPg*)
    std::vector<float> basisT_hi, basisT_lo, basis_hi, basis_lo, w_hi, w_lo, wT_hi, wT_lo;   // tf32 hi/lo operands of the tcgen05 kernels
    ModelView view;      // integer tables filled in; pointers left null
    bool has_prior = false;
};

// Returns an empty string on success, otherwise the validation error.
std::string build_host_model(const smplb200_model_desc& d, HostModel& out);

}  // namespace smplb200
