// Subset posterior math shared by the model and DAA kernels.
//   poe_fusion + poe            BaseMMVae.py:109-122, divergence_measures/mm_div.py:13-20
//   moe_fusion / selection      BaseMMVae.py:96-106, utils/utils.py:63-85
//   fusion_condition_*          BaseMMVae.py:125-134
#pragma once
#include "mopoe_common.cuh"

namespace mopoe {

// moe and jsd fuse the members of a subset by selection (BaseMMVae.py:46-54) and put the unimodal experts in the mixture
__device__ __forceinline__ bool moe_like(const ModelView& mv) { return mv.method == MOPOE_METHOD_MOE || mv.method == MOPOE_METHOD_JSD; }
// jsd: the last mixture component is the prior N(0, I) (BaseMMVae.py:217-223), which is no subset of the table
__device__ __forceinline__ bool prior_component(const ModelView& mv, const mopoe_batch_desc& b, int k) {
  return mv.method == MOPOE_METHOD_JSD && k == b.n_mix - 1;
}

// posterior of subset s at one (row, latent) element from the experts (mu_e, lv_e)
struct SubsetEval {
  float mu, lv, sumT;
  int sel;  // moe multi-member: chosen member
};

__device__ __forceinline__ SubsetEval eval_subset(const ModelView& mv, const mopoe_batch_desc& b, int s,
                                                  int n, const float* mu_e, const float* lv_e) {
  SubsetEval r;
  const int nm = mv.sub.n_members[s];
  r.sel = 0; r.sumT = 1.f;
  if (moe_like(mv)) {  // moe_fusion -> mixture_component_selection
    for (int i = 0; i < nm; ++i)
      if (n >= b.moe_bounds[nm][i] && n < b.moe_bounds[nm][i + 1]) r.sel = i;
    const int m = mv.sub.members[s][r.sel];
    r.mu = mu_e[m]; r.lv = lv_e[m];
  } else {                               // poe_fusion + poe (mm_div.py:13-20)
    float sT = 0.f, sMT = 0.f;
    for (int i = 0; i < nm; ++i) {
      const int m = mv.sub.members[s][i];
      const float T = 1.f / (expf(lv_e[m]) + MOPOE_POE_EPS);
      sT += T; sMT += mu_e[m] * T;
    }
    if (mv.method == MOPOE_METHOD_POE || nm == mv.M) {  // prior expert N(0, I)
      sT += 1.f / (1.f + MOPOE_POE_EPS);
    }
    r.sumT = sT;
    r.mu = sMT / sT;
    r.lv = logf(1.f / sT);
  }
  return r;
}

__device__ __forceinline__ bool in_mixture(const ModelView& mv, const mopoe_batch_desc& b, int s) {
  if (moe_like(mv)) return mv.sub.n_members[s] == 1;                          // fusion_condition_moe
  if (mv.method == MOPOE_METHOD_POE) return mv.sub.mask[s] == b.present_mask;  // fusion_condition_poe
  return true;                                                                // fusion_condition_joint
}


}  // namespace mopoe
