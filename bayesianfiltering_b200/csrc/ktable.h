// ktable.h -- how the host translation unit (bfilt_api.cu) reaches the kernels compiled in the other translation units.
//
// One nvcc process over every model template takes six minutes (ptxas is serial per translation unit, and
// --split-compile makes register allocation vary from build to build), so the model-templated kernels are compiled as
// slices of the registry in separate translation units (tu_kernels.cu, see csrc/build.sh) and handed to the host code
// as plain function addresses: the host launches them with cudaLaunchKernel / cudaLaunchCooperativeKernel.
#pragma once
#include "common.cuh"

struct BfKernelSet {
  const void* fused[4];       // k_pf_fused<M, ...>: general / sharded / lean / sharded lean   args (ModelConst, PfFusedArgs)
  int fused_warps;            // warps per CTA of the fused kernels of this model
  const void* resident[2];    // k_pf_resident<M, 8 / 32>                       args (ModelConst, PfResidentArgs)
  const void* gsf_cta;        // k_gsf<M, TV = false>                           args (ModelConst, GsfArgs)
  const void* gsf_cta_tv;     // k_gsf<M, TV = true>: time-varying Q_t / R_t
  const void* gsf_warp[4];    // k_gsf_warp<M, 1 / 2 / 4 / 8>                   args (ModelConst, GsfArgs, int B)
  const void* ugsf;           // k_ugsf<M>: unscented Gaussian-sum filter       args (ModelConst, GsfArgs, UkfArgs)
  const void* agsf;           // k_agsf<M>: augmented Gaussian-sum filter       args (ModelConst, AgsfArgs)
  const void* model_eval;     // k_model_eval<M>
  const void* ssm_sample;     // k_ssm_sample<M>
};

enum { BFK_FUSED = 1, BFK_RESIDENT = 2, BFK_GSF = 3, BFK_MISC = 4, BFK_LARGE = 5 };

// Each slice exports  int bfk_fill_<kind>_<slice>(const ModelConst&, BfKernelSet*)  : fills its fields and returns 0,
// or returns -3 when the model is not in that slice.
typedef int (*bfk_fill_fn)(const ModelConst&, BfKernelSet*);

// k_gsf_large<40> (dxt = 40) or the generic k_gsf_large<0>
const void* bfk_gsf_large(int dxt);
