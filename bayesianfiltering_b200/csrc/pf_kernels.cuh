// pf_kernels.cuh -- bootstrap particle filter kernels (sm_100a).
//
// Restates the body of gaussfiltax/inference.py:1330-1366 (`_step` of bootstrap_particle_filter) and
// gaussfiltax/utils.py:207-214 (`_resample`) as two device paths:
//
//  STREAMED  (large N, state in HBM, 5 launches per step, no host sync -- the ESS branch is a device flag)
//    k_propagate   gather x[a_i] (the previous step's resampling, fused), x~ = f(x,q,u), ll = log p(y|x~),
//                  per-CTA online (max, sum e, sum e^2, sum e x) partials; the last CTA to finish folds them
//                  into the step's stats: max, Z, log-lik increment, ESS, resample flag, weighted mean
//    k_weights     w = exp(ll-max) wprev / Z ; no-resample: store w ; resample: tile tree sums of w
//    k_scan_tiles  canonical exclusive scan of the tile sums (one CTA)
//    k_downsweep   recompute w, canonical exclusive scan inside each tile -> E[0..N]  (cdf[j] = E[j+1])
//    k_search      r_i from u_i, fixed-bisection lower bound in E -> ancestor a_i
//  RESIDENT  (N small enough for shared memory) one CTA per independent trial runs ALL T steps with the
//    particles, weights, cdf and ancestors resident in shared memory: `k_pf_resident`.
//
// Particle state is particle-major [N][DX] (the reference's own [N,dx] layout): one 128-bit load per
// particle for DX=4, which is what the random ancestor gather wants (one 32-B sector per particle instead
// of DX sectors for a structure-of-arrays layout).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "scan.cuh"

#define PF_TILE_WARPS 32
#define PF_TILE (PF_TILE_WARPS * 128)      // 4096 leaves per scan tile
#define PF_PROP_THREADS 256

struct StepStats {
  float mx, Z, loglik, ess;
  int flag;        // this step resampled: the next step gathers through anc[] and sees uniform weights
  int w_uniform;   // weights after this step are exactly 1/N
  int pad0, pad1;
  float mean[BF_MAXD];
};

struct PfPartial { float m, s, s2; float sx[BF_MAXD]; };

// ---------------------------------------------------------------- vectorised particle access
template <int DX>
__device__ __forceinline__ void load_state(const float* __restrict__ base, size_t i, float* x) {
  if constexpr (DX % 4 == 0) {
    const float4* p = reinterpret_cast<const float4*>(base) + i * (DX / 4);
#pragma unroll
    for (int k = 0; k < DX / 4; ++k) { float4 v = p[k]; x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w; }
  } else if constexpr (DX % 2 == 0) {
    const float2* p = reinterpret_cast<const float2*>(base) + i * (DX / 2);
#pragma unroll
    for (int k = 0; k < DX / 2; ++k) { float2 v = p[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
  } else {
#pragma unroll
    for (int k = 0; k < DX; ++k) x[k] = base[i * DX + k];
  }
}
template <int DX>
__device__ __forceinline__ void store_state(float* __restrict__ base, size_t i, const float* x) {
  if constexpr (DX % 4 == 0) {
    float4* p = reinterpret_cast<float4*>(base) + i * (DX / 4);
#pragma unroll
    for (int k = 0; k < DX / 4; ++k) p[k] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
  } else if constexpr (DX % 2 == 0) {
    float2* p = reinterpret_cast<float2*>(base) + i * (DX / 2);
#pragma unroll
    for (int k = 0; k < DX / 2; ++k) p[k] = make_float2(x[2 * k], x[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < DX; ++k) base[i * DX + k] = x[k];
  }
}

// ---------------------------------------------------------------- online (max, sum) accumulator
template <int DX>
struct OnlineAcc {
  float m, s, s2, sx[DX];
  __device__ __forceinline__ void init() {
    m = -INFINITY; s = 0.0f; s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] = 0.0f;
  }
  __device__ __forceinline__ void rescale(float nm) {   // nm >= m
    float sc = (m == -INFINITY) ? 0.0f : __expf(m - nm);
    s *= sc; s2 *= sc * sc;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] *= sc;
    m = nm;
  }
  __device__ __forceinline__ void add(float ll, float wprev, const float* x) {
    if (ll > m) rescale(ll);
    // statistics only (sum, ESS, mean): the fast exponential's ~1e-7 relative error is far inside their 1e-5 bar;
    // the weights themselves are computed with expf in norm_weight*
    float e = (ll == -INFINITY) ? 0.0f : __expf(ll - m) * wprev;
    s += e; s2 += e * e;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] += e * x[k];
  }
  __device__ __forceinline__ void merge(float om, float os, float os2, const float* osx) {
    float nm = fmaxf(m, om);
    float a = (m == -INFINITY) ? 0.0f : expf(m - nm);
    float b = (om == -INFINITY) ? 0.0f : expf(om - nm);
    if (om != om || m != m) nm = NAN;             // NaN log-lik poisons the step like jnp.max does
    s = s * a + os * b; s2 = s2 * a * a + os2 * b * b;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] = sx[k] * a + osx[k] * b;
    m = nm;
  }
  __device__ __forceinline__ void warp_reduce() {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float om = __shfl_xor_sync(BF_FULL, m, o), os = __shfl_xor_sync(BF_FULL, s, o), os2 = __shfl_xor_sync(BF_FULL, s2, o);
      float osx[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) osx[k] = __shfl_xor_sync(BF_FULL, sx[k], o);
      merge(om, os, os2, osx);
    }
  }
};

// parameters of one streamed step (device pointers already offset to the current trial)
struct PfStepArgs {
  int N;                    // particles
  int t;                    // step index inside this call
  int t_global;             // t + t_offset (Philox counter)
  int trial;                // trial id (Philox counter)
  float ess_threshold;
  int resampler;
  unsigned long long seed;
  const float* y;           // [DY] emission of this step
  const float* u_in;        // inputs + t, or NULL
  const float* x_prev;      // [N,DX]
  float* x_cur;             // [N,DX]
  float* W;                 // [N] carried weights (valid when !w_uniform)
  float* LL;                // [N + pad]
  float* E;                 // [ntiles*PF_TILE] exclusive canonical cdf, E[N] = total
  int* anc;                 // [N]
  float* tilesum;           // [ntiles]
  float* Etile;             // [ntiles+1]
  PfPartial* partials;      // [grid of k_propagate]
  unsigned int* counter;    // last-CTA ticket
  const StepStats* prev;    // stats of the previous step boundary
  StepStats* cur;           // stats of this step
  const float* noise;       // [N,DQ] coloured noise of this step, or NULL (Philox)
  const float* res_u;       // multinomial [N] / systematic [1], or NULL (Philox)
  // optional outputs of this step (already offset), NULL to skip
  float* out_particles; float* out_weights; float* out_propagated; float* out_norm_weights; int* out_ancestors;
  float* out_mean; float* out_loglik; float* out_ess; int* out_resampled;
};

// ---------------------------------------------------------------- k_propagate
template <class M>
__global__ void __launch_bounds__(PF_PROP_THREADS)
k_propagate(const __grid_constant__ ModelConst mc, const __grid_constant__ PfStepArgs a) {
  constexpr int DX = M::DX, DQ = M::DQ, DY = M::DY;
  __shared__ PfPartial s_part[PF_PROP_THREADS / 32];
  __shared__ bool s_last;
  const int N = a.N;
  const bool gather = a.prev->flag != 0;
  const bool w_uni = a.prev->w_uniform != 0;
  const float invN = 1.0f / (float)N;
  const float u = a.u_in ? a.u_in[0] : 0.0f;
  float y[DY];
#pragma unroll
  for (int k = 0; k < DY; ++k) y[k] = a.y[k];

  OnlineAcc<DX> acc; acc.init();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const int src = gather ? a.anc[i] : i;
    float x[DX], q[DQ], xt[DX];
    load_state<DX>(a.x_prev, (size_t)src, x);
    if (a.noise) {
#pragma unroll
      for (int k = 0; k < DQ; ++k) q[k] = a.noise[(size_t)i * DQ + k];
    } else {
      float z[DQ];
      philox_normals<DQ>(a.seed, BF_STREAM_NOISE, (uint32_t)a.t_global, (uint32_t)a.trial, (uint64_t)i, z);
#pragma unroll
      for (int r = 0; r < DQ; ++r) {          // q = q0 + L_Q z  (models.py:83, TFP MVN sample)
        float s = mc.q0[r];
#pragma unroll
        for (int c = 0; c <= r; ++c) s += mc.LQ[r * DQ + c] * z[c];
        q[r] = s;
      }
    }
    M::f(x, q, u, mc, xt);
    store_state<DX>(a.x_cur, (size_t)i, xt);
    if (a.out_propagated) store_state<DX>(a.out_propagated, (size_t)i, xt);
    const float ll = emission_log_prob<M>(xt, y, u, mc);
    a.LL[i] = ll;
    const float wp = w_uni ? invN : a.W[i];
    acc.add(ll, wp, xt);
  }
  // CTA reduction
  acc.warp_reduce();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_part[warp].m = acc.m; s_part[warp].s = acc.s; s_part[warp].s2 = acc.s2;
#pragma unroll
    for (int k = 0; k < DX; ++k) s_part[warp].sx[k] = acc.sx[k];
  }
  __syncthreads();
  if (warp == 0) {
    acc.init();
    if (lane < PF_PROP_THREADS / 32) acc.merge(s_part[lane].m, s_part[lane].s, s_part[lane].s2, s_part[lane].sx);
    acc.warp_reduce();
    if (lane == 0) {
      PfPartial* p = &a.partials[blockIdx.x];
      p->m = acc.m; p->s = acc.s; p->s2 = acc.s2;
#pragma unroll
      for (int k = 0; k < DX; ++k) p->sx[k] = acc.sx[k];
      __threadfence();
      unsigned int ticket = atomicAdd(a.counter, 1u);
      s_last = (ticket == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA: fold all partials in a fixed order (deterministic for a given grid)
  __threadfence();
  acc.init();
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    const volatile PfPartial* p = &a.partials[b];
    float sx[DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] = p->sx[k];
    acc.merge(p->m, p->s, p->s2, sx);
  }
  acc.warp_reduce();
  if (lane == 0) {
    s_part[warp].m = acc.m; s_part[warp].s = acc.s; s_part[warp].s2 = acc.s2;
#pragma unroll
    for (int k = 0; k < DX; ++k) s_part[warp].sx[k] = acc.sx[k];
  }
  __syncthreads();
  if (warp == 0) {
    acc.init();
    if (lane < PF_PROP_THREADS / 32) acc.merge(s_part[lane].m, s_part[lane].s, s_part[lane].s2, s_part[lane].sx);
    acc.warp_reduce();
    if (lane == 0) {
      StepStats st;
      st.mx = acc.m; st.Z = acc.s;
      st.loglik = acc.m + logf(acc.s);                       // log p(y_t | y_1:t-1)
      st.ess = acc.s * acc.s / acc.s2;                       // 1 / sum (e/Z)^2   (inference.py:1356)
      st.flag = (st.ess < a.ess_threshold * (float)N) ? 1 : 0;
      st.w_uniform = st.flag;
      st.pad0 = st.pad1 = 0;
#pragma unroll
      for (int k = 0; k < BF_MAXD; ++k) st.mean[k] = (k < DX) ? acc.sx[k] / acc.s : 0.0f;
      *a.cur = st;
      if (a.out_loglik) a.out_loglik[0] = st.loglik;
      if (a.out_ess) a.out_ess[0] = st.ess;
      if (a.out_resampled) a.out_resampled[0] = st.flag;
      if (a.out_mean) {
#pragma unroll
        for (int k = 0; k < DX; ++k) a.out_mean[k] = st.mean[k];
      }
      *a.counter = 0u;
    }
  }
}

// normalised weight of particle i exactly as inference.py:1350-1353: (exp(ll - max) * w_prev) / sum
__device__ __forceinline__ float norm_weight(float ll, float wprev, float mx, float Z) {
  return __fdiv_rn(__fmul_rn(expf(__fsub_rn(ll, mx)), wprev), Z);
}
// same with the reciprocal of the sum hoisted (<= 1.5 ulp from the division; used consistently in every phase)
__device__ __forceinline__ float norm_weight_rcp(float ll, float wprev, float mx, float rZ) {
  return __fmul_rn(__fmul_rn(expf(__fsub_rn(ll, mx)), wprev), rZ);
}

// weights of the 4 leaves of this thread in tile `tile` (zero past N)
__device__ __forceinline__ void tile_weights(const PfStepArgs& a, int tile, float mx, float Z, bool w_uni, float v[4]) {
  const int base = tile * PF_TILE + threadIdx.x * 4;
  const float invN = 1.0f / (float)a.N;
  const float4 l4 = *reinterpret_cast<const float4*>(a.LL + base);     // LL is padded to a whole tile
  float wp[4] = {invN, invN, invN, invN};
  if (!w_uni) {
#pragma unroll
    for (int k = 0; k < 4; ++k) wp[k] = (base + k < a.N) ? a.W[base + k] : 0.0f;
  }
  const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (base + k < a.N) ? norm_weight(ll[k], wp[k], mx, Z) : 0.0f;
}

// ---------------------------------------------------------------- k_weights
__global__ void __launch_bounds__(PF_TILE_WARPS * 32)
k_weights(const __grid_constant__ PfStepArgs a) {
  __shared__ float s_warp[PF_TILE_WARPS];
  const StepStats st = *a.cur;
  const bool w_uni = a.prev->w_uniform != 0;
  float v[4];
  tile_weights(a, blockIdx.x, st.mx, st.Z, w_uni, v);
  const int base = blockIdx.x * PF_TILE + threadIdx.x * 4;
  if (a.out_norm_weights) {
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < a.N) a.out_norm_weights[base + k] = v[k];
  }
  if (!st.flag) {                      // identity branch of lax.cond (inference.py:1357): carry the weights
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < a.N) {
      a.W[base + k] = v[k];
      if (a.out_weights) a.out_weights[base + k] = v[k];
    }
    return;
  }
  if (a.out_weights) {                 // utils.py:212: weights = 1/N after resampling
    const float invN = 1.0f / (float)a.N;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < a.N) a.out_weights[base + k] = invN;
  }
  const float t = tile_tree_sum<PF_TILE_WARPS>(v, s_warp);
  if (threadIdx.x == 0) a.tilesum[blockIdx.x] = t;
}

// ---------------------------------------------------------------- k_scan_tiles (one CTA)
__global__ void __launch_bounds__(PF_TILE_WARPS * 32)
k_scan_tiles(const float* tilesum, float* Etile, int ntiles, const StepStats* cur) {
  extern __shared__ float s_scan[];
  if (cur && !cur->flag) return;
  cta_exclusive_scan<PF_TILE_WARPS>(tilesum, Etile, ntiles, s_scan);
}

// ---------------------------------------------------------------- k_downsweep
__global__ void __launch_bounds__(PF_TILE_WARPS * 32)
k_downsweep(const __grid_constant__ PfStepArgs a) {
  __shared__ float s_warp[PF_TILE_WARPS];
  const StepStats st = *a.cur;
  if (!st.flag) return;
  const bool w_uni = a.prev->w_uniform != 0;
  float v[4], e[4];
  tile_weights(a, blockIdx.x, st.mx, st.Z, w_uni, v);
  tile_exclusive_scan<PF_TILE_WARPS>(v, a.Etile[blockIdx.x], e, s_warp);
  *reinterpret_cast<float4*>(a.E + (size_t)blockIdx.x * PF_TILE + threadIdx.x * 4) = make_float4(e[0], e[1], e[2], e[3]);
}

// resampling position of output slot i (oracle/scan.py multinomial_positions / systematic_positions)
__device__ __forceinline__ float resample_position(int resampler, float total, float u, int i, int N) {
  if (resampler == 1) return __fmul_rn(total, __fdiv_rn(__fadd_rn((float)i, u), (float)N));
  return __fmul_rn(total, __fsub_rn(1.0f, u));
}

// ---------------------------------------------------------------- k_search
__global__ void __launch_bounds__(256)
k_search(const __grid_constant__ PfStepArgs a) {
  const bool flag = a.cur->flag != 0;
  const int N = a.N;
  if (!flag) {
    if (a.out_ancestors)
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) a.out_ancestors[i] = i;
    return;
  }
  const float total = a.E[N];
  const int levels = bisect_levels(N);
  float u0 = 0.0f;
  if (a.resampler == 1)
    u0 = a.res_u ? a.res_u[0] : philox_uniform(a.seed, BF_STREAM_RESAMPLE, (uint32_t)a.t_global, (uint32_t)a.trial, 0ull);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    float u = u0;
    if (a.resampler == 0)
      u = a.res_u ? a.res_u[i] : philox_uniform(a.seed, BF_STREAM_RESAMPLE, (uint32_t)a.t_global, (uint32_t)a.trial, (uint64_t)i);
    const float r = resample_position(a.resampler, total, u, i, N);
    const int anc = bisect_left_exclusive(a.E, N, levels, r);
    a.anc[i] = anc;
    if (a.out_ancestors) a.out_ancestors[i] = anc;
  }
}

// ---------------------------------------------------------------- k_emit: post-resample particles of this step
template <int DX>
__global__ void __launch_bounds__(256)
k_emit(const float* __restrict__ x_cur, const int* __restrict__ anc, const StepStats* cur, int N, float* __restrict__ out) {
  const bool flag = cur->flag != 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    float x[DX];
    load_state<DX>(x_cur, (size_t)(flag ? anc[i] : i), x);
    store_state<DX>(out, (size_t)i, x);
  }
}

// materialise the carry after the last step: x_out[i] = x_cur[anc[i]] (or copy), weights 1/N or W
template <int DX>
__global__ void __launch_bounds__(256)
k_finalize(const float* __restrict__ x_cur, const int* __restrict__ anc, StepStats* cur, int N, float* __restrict__ x_out,
           float* __restrict__ W, float* __restrict__ final_x, float* __restrict__ final_w) {
  const bool flag = cur->flag != 0;
  const bool w_uni = cur->w_uniform != 0;
  const float invN = 1.0f / (float)N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    float x[DX];
    load_state<DX>(x_cur, (size_t)(flag ? anc[i] : i), x);
    store_state<DX>(x_out, (size_t)i, x);
    if (final_x) store_state<DX>(final_x, (size_t)i, x);
    if (final_w) final_w[i] = w_uni ? invN : W[i];
  }
}

// after k_finalize the carry is materialised: clear the gather flag for a resumed run
__global__ void k_clear_gather(StepStats* st) { st->flag = 0; }

// initial particles ~ N(m0, P0) from Philox (inference.py:1372-1373 with our own stream)
template <int DX>
__global__ void __launch_bounds__(256)
k_init_particles(const __grid_constant__ ModelConst mc, float* __restrict__ x, int N, unsigned long long seed, int trial, int goff) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    float z[DX], o[DX];
    philox_normals<DX>(seed, BF_STREAM_INIT, 0u, (uint32_t)trial, (uint64_t)(goff + i), z);   // goff: shard of a sharded filter
#pragma unroll
    for (int r = 0; r < DX; ++r) {
      float s = mc.m0[r];
#pragma unroll
      for (int c = 0; c <= r; ++c) s += mc.LP0[r * DX + c] * z[c];
      o[r] = s;
    }
    store_state<DX>(x, (size_t)i, o);
  }
}

__global__ void k_init_stats(StepStats* st, int w_uniform) {
  st->mx = 0.0f; st->Z = 1.0f; st->loglik = 0.0f; st->ess = 0.0f; st->flag = 0; st->w_uniform = w_uniform;
  st->pad0 = st->pad1 = 0;
  for (int k = 0; k < BF_MAXD; ++k) st->mean[k] = 0.0f;
}

// ---------------------------------------------------------------- stand-alone scan pieces (bf_cumsum / bf_resample)
__global__ void __launch_bounds__(PF_TILE_WARPS * 32)
k_raw_tilesum(const float* __restrict__ w, int n, float* tilesum) {
  __shared__ float s_warp[PF_TILE_WARPS];
  float v[4];
  const int base = blockIdx.x * PF_TILE + threadIdx.x * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (base + k < n) ? w[base + k] : 0.0f;
  const float t = tile_tree_sum<PF_TILE_WARPS>(v, s_warp);
  if (threadIdx.x == 0) tilesum[blockIdx.x] = t;
}
__global__ void __launch_bounds__(PF_TILE_WARPS * 32)
k_raw_downsweep(const float* __restrict__ w, int n, const float* Etile, float* E) {
  __shared__ float s_warp[PF_TILE_WARPS];
  float v[4], e[4];
  const int base = blockIdx.x * PF_TILE + threadIdx.x * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (base + k < n) ? w[base + k] : 0.0f;
  tile_exclusive_scan<PF_TILE_WARPS>(v, Etile[blockIdx.x], e, s_warp);
  *reinterpret_cast<float4*>(E + (size_t)base) = make_float4(e[0], e[1], e[2], e[3]);
}
__global__ void k_shift_cdf(const float* __restrict__ E, float* __restrict__ cdf, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cdf[i] = E[i + 1];
}
__global__ void k_raw_search(const float* __restrict__ E, const float* __restrict__ u, int* __restrict__ anc, int n, int resampler) {
  const float total = E[n];
  const int levels = bisect_levels(n);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float uu = (resampler == 1) ? u[0] : u[i];
    const float r = resample_position(resampler, total, uu, i, n);
    anc[i] = bisect_left_exclusive(E, n, levels, r);
  }
}

// ---------------------------------------------------------------- model evaluation + data generator
// what: 0 f(x,q,u) [n,DX] ; 1 h(x,r,u) [n,DY] ; 2 Fx [n,DX,DX] ; 3 Fq [n,DX,DQ] ; 4 Hx [n,DY,DX] ;
//       5 emission log-prob of y (aux = y [DY], shared) [n]
template <class M>
__global__ void k_model_eval(const __grid_constant__ ModelConst mc, int what, int n, const float* __restrict__ x,
                             const float* __restrict__ aux, float u, float* __restrict__ out) {
  constexpr int DX = M::DX, DQ = M::DQ, DY = M::DY, DR = M::DR;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float xi[DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) xi[k] = x[(size_t)i * DX + k];
    if (what == 0) {
      float q[DQ], o[DX];
#pragma unroll
      for (int k = 0; k < DQ; ++k) q[k] = aux ? aux[(size_t)i * DQ + k] : mc.q0[k];
      M::f(xi, q, u, mc, o);
#pragma unroll
      for (int k = 0; k < DX; ++k) out[(size_t)i * DX + k] = o[k];
    } else if (what == 1) {
      float r[DR], o[DY];
#pragma unroll
      for (int k = 0; k < DR; ++k) r[k] = aux ? aux[(size_t)i * DR + k] : mc.r0[k];
      M::h(xi, r, u, mc, o);
#pragma unroll
      for (int k = 0; k < DY; ++k) out[(size_t)i * DY + k] = o[k];
    } else if (what == 2) {
      float J[DX * DX];
      M::Fx(xi, mc.q0, u, mc, J);
#pragma unroll
      for (int k = 0; k < DX * DX; ++k) out[(size_t)i * DX * DX + k] = J[k];
    } else if (what == 3) {
      float J[DX * DQ];
      M::Fq(xi, mc.q0, u, mc, J);
#pragma unroll
      for (int k = 0; k < DX * DQ; ++k) out[(size_t)i * DX * DQ + k] = J[k];
    } else if (what == 4) {
      float J[DY * DX];
      M::Hx(xi, mc.r0, u, mc, J);
#pragma unroll
      for (int k = 0; k < DY * DX; ++k) out[(size_t)i * DY * DX + k] = J[k];
    } else {
      float y[DY];
#pragma unroll
      for (int k = 0; k < DY; ++k) y[k] = aux[k];
      out[i] = emission_log_prob<M>(xi, y, u, mc);
    }
  }
}

// NonlinearSSM.sample (gaussfiltax/models.py:240-289): z_0 ~ N(m0,P0), y_0 = h(z_0, r_0, u_0); then
// z_t = f(z_{t-1}, q_t, u_t), y_t = h(z_t, r_t, u_t).  One thread per independent trial (time is
// sequential).  Noise: supplied (already coloured) or Philox.
template <class M>
__global__ void k_ssm_sample(const __grid_constant__ ModelConst mc, int B, int T, const float* __restrict__ inputs,
                             unsigned long long seed, const float* __restrict__ z0, const float* __restrict__ qn,
                             const float* __restrict__ rn, float* __restrict__ states, float* __restrict__ emissions) {
  constexpr int DX = M::DX, DQ = M::DQ, DY = M::DY, DR = M::DR;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float z[DX];
  if (z0) {
#pragma unroll
    for (int k = 0; k < DX; ++k) z[k] = z0[(size_t)b * DX + k];
  } else {
    float n[DX];
    philox_normals<DX>(seed, 11u, 0u, (uint32_t)b, 0ull, n);
#pragma unroll
    for (int r = 0; r < DX; ++r) {
      float s = mc.m0[r];
#pragma unroll
      for (int c = 0; c <= r; ++c) s += mc.LP0[r * DX + c] * n[c];
      z[r] = s;
    }
  }
  for (int t = 0; t < T; ++t) {
    const float u = inputs ? inputs[t] : 0.0f;
    const size_t bt = (size_t)b * T + t;
    if (t > 0) {
      float q[DQ], zn[DX];
      if (qn) {
#pragma unroll
        for (int k = 0; k < DQ; ++k) q[k] = qn[bt * DQ + k];
      } else {
        float n[DQ];
        philox_normals<DQ>(seed, 12u, (uint32_t)t, (uint32_t)b, 0ull, n);
#pragma unroll
        for (int r = 0; r < DQ; ++r) {
          float s = mc.q0[r];
#pragma unroll
          for (int c = 0; c <= r; ++c) s += mc.LQ[r * DQ + c] * n[c];
          q[r] = s;
        }
      }
      M::f(z, q, u, mc, zn);
#pragma unroll
      for (int k = 0; k < DX; ++k) z[k] = zn[k];
    }
    float r[DR], y[DY];
    if (rn) {
#pragma unroll
      for (int k = 0; k < DR; ++k) r[k] = rn[bt * DR + k];
    } else {
      float n[DR];
      philox_normals<DR>(seed, 13u, (uint32_t)t, (uint32_t)b, 0ull, n);
#pragma unroll
      for (int rr = 0; rr < DR; ++rr) {
        float s = mc.r0[rr];
#pragma unroll
        for (int c = 0; c <= rr; ++c) s += mc.LR[rr * DR + c] * n[c];
        r[rr] = s;
      }
    }
    M::h(z, r, u, mc, y);
#pragma unroll
    for (int k = 0; k < DX; ++k) states[bt * DX + k] = z[k];
#pragma unroll
    for (int k = 0; k < DY; ++k) emissions[bt * DY + k] = y[k];
  }
}
