// tu_kernels.cu -- one slice of the model-templated kernels.  Compiled several times by csrc/build.sh with
//   -DBF_TU_KIND=<BFK_*>  -DBF_NGROUPS=<slices of this kind>  -DBF_GROUP=<this slice>
// and exports the addresses of the kernels it instantiated (ktable.h).  No host logic lives here.
#include "ktable.h"
#include "models.cuh"
#include "scan.cuh"
#include "pf_kernels.cuh"

#define BFK_NAME2(kind, g) bfk_fill_##kind##_##g
#define BFK_NAME(kind, g) BFK_NAME2(kind, g)

#if BF_TU_KIND == 1  // ------------------------------------------------------------------ fused particle filter
#include "pf_fused.cuh"
namespace {
struct Fill {
  BfKernelSet* k;
  template <class M> int run() {
    // small states: ONE CTA of 32 warps per SM at 64 registers per thread.  (Two CTAs of 16 warps do not share an SM
    // evenly -- the scheduler favours the older one, which then idles ~20 us per phase at the grid barrier while the
    // younger one finishes with half the SM's warps.)  Large states (dx > 4) need ~128 registers per thread to keep a
    // particle, its noise and the statistics out of local memory: one CTA of 16 warps per SM.
    constexpr int NW = (M::DX > 4) ? PF_FUSED_WARPS_LARGE : PF_FUSED_WARPS;
    k->fused_warps = NW;
    k->fused[0] = (const void*)k_pf_fused<M, NW, 1, false, false>;
    k->fused[1] = (const void*)k_pf_fused<M, NW, 1, true, false>;
    k->fused[2] = (const void*)k_pf_fused<M, NW, 1, false, true>;
    k->fused[3] = (const void*)k_pf_fused<M, NW, 1, true, true>;
    return 0;
  }
};
}  // namespace
int BFK_NAME(fused, BF_GROUP)(const ModelConst& mc, BfKernelSet* k) { Fill f{k}; return bf_dispatch_model(mc, f); }

#elif BF_TU_KIND == 2  // ---------------------------------------------------------------- resident particle filter
#include "pf_resident.cuh"
namespace {
struct Fill {
  BfKernelSet* k;
  template <class M> int run() {
    k->resident[0] = (const void*)k_pf_resident<M, 8>;
    k->resident[1] = (const void*)k_pf_resident<M, 32>;
    return 0;
  }
};
}  // namespace
int BFK_NAME(resident, BF_GROUP)(const ModelConst& mc, BfKernelSet* k) { Fill f{k}; return bf_dispatch_model(mc, f); }

#elif BF_TU_KIND == 3  // ---------------------------------------------------------------- Gaussian-sum filter
#include "ugsf_kernels.cuh"
#include "agsf_kernels.cuh"
namespace {
struct Fill {
  BfKernelSet* k;
  template <class M> int run() {
    k->gsf_cta = (const void*)k_gsf<M, false>;
    k->gsf_cta_tv = (const void*)k_gsf<M, true>;
    k->gsf_warp[0] = (const void*)k_gsf_warp<M, 1>;
    k->gsf_warp[1] = (const void*)k_gsf_warp<M, 2>;
    k->gsf_warp[2] = (const void*)k_gsf_warp<M, 4>;
    k->gsf_warp[3] = (const void*)k_gsf_warp<M, 8>;
    if constexpr (M::DX <= 4) k->agsf = (const void*)k_agsf<M>;
    if constexpr (M::DX <= 4) k->ugsf = (const void*)k_ugsf<M>;      // sigma points of the augmented state live in registers
    return 0;
  }
};
}  // namespace
int BFK_NAME(gsf, BF_GROUP)(const ModelConst& mc, BfKernelSet* k) { Fill f{k}; return bf_dispatch_model(mc, f); }

#elif BF_TU_KIND == 4  // ---------------------------------------------------------------- model evaluation, data generator
namespace {
struct Fill {
  BfKernelSet* k;
  template <class M> int run() {
    k->model_eval = (const void*)k_model_eval<M>;
    k->ssm_sample = (const void*)k_ssm_sample<M>;
    return 0;
  }
};
}  // namespace
int BFK_NAME(misc, BF_GROUP)(const ModelConst& mc, BfKernelSet* k) { Fill f{k}; return bf_dispatch_model(mc, f); }

#elif BF_TU_KIND == 5  // ---------------------------------------------------------------- large-state Gaussian-sum filter
#include "gsf_large.cuh"
const void* bfk_gsf_large(int dxt) { return dxt == 40 ? (const void*)k_gsf_large<40> : (const void*)k_gsf_large<0>; }

#else
#error "BF_TU_KIND must be 1..5"
#endif
