// pf_kernels.cuh -- shared pieces of the bootstrap particle filter kernels (sm_100a).
//
// The filter itself -- gaussfiltax/inference.py:1330-1366 (`_step` of bootstrap_particle_filter) and
// gaussfiltax/utils.py:207-214 (`_resample`) -- lives in two kernels:
//   pf_fused.cuh     large N, state in HBM: one persistent cooperative kernel runs all T steps
//   pf_resident.cuh  N small enough for shared memory: one CTA per independent trial runs all T steps
// This header holds what both share (step statistics, the online (max, sum) accumulator, the weight formula),
// the stand-alone scan / search kernels behind bf_cumsum / bf_resample, the Philox initial-particle kernel and the
// model evaluation / data generator kernels.
//
// Particle state is particle-major [N][DX] (the reference's own [N,dx] layout): one 128-bit load per
// particle for DX=4, which is what the random ancestor gather wants (one 32-B sector per particle instead
// of DX sectors for a structure-of-arrays layout).
//
// Non-template kernels are compiled only into the translation unit that defines BF_TU_MAIN (bfilt_api.cu); the model
// templates are instantiated by the kernel translation units (tu_kernels.cu), several nvcc processes in parallel.
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

struct __align__(16) PfPartial { float m, s, s2, pad; float sx[BF_MAXD]; };     // 48 bytes: three 128-bit accesses
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
    float sc = (m == -INFINITY) ? 0.0f : mufu_ex2((m - nm) * 1.44269504088896340736f);
    s *= sc; s2 *= sc * sc;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] *= sc;
    m = nm;
  }
  // P particles at once: ONE test against the running maximum (and at most one rescale) for the group -- ptxas turns the
  // rescale into predicated code, so it is paid whether taken or not.  Statistics only (sum, ESS, mean): MUFU.EX2's ~1e-7
  // relative error is far inside their 1e-5 bar; the weights themselves use expf (norm_weight*).  The difference
  // ll - m is formed first (exact for nearby values) and only then scaled by log2(e): log-likelihoods of size 1e3
  // (R = 25e-6) would otherwise lose the digits the exponential needs.
  template <int P>
  __device__ __forceinline__ void add_group(const float* ll, const float* wprev, const float (*x)[DX]) {
    float gm = ll[0];
#pragma unroll
    for (int p = 1; p < P; ++p) gm = fmaxf(gm, ll[p]);
    if (gm > m) rescale(gm);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float e = (ll[p] == -INFINITY) ? 0.0f : mufu_ex2((ll[p] - m) * 1.44269504088896340736f) * wprev[p];
      s += e; s2 = fmaf(e, e, s2);
#pragma unroll
      for (int k = 0; k < DX; ++k) sx[k] = fmaf(e, x[p][k], sx[k]);
    }
  }
  __device__ __forceinline__ void add(float ll, float wprev, const float* x) {
    float xx[1][DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) xx[0][k] = x[k];
    add_group<1>(&ll, &wprev, xx);
  }
  __device__ __forceinline__ void merge(float om, float os, float os2, const float* osx) {
    float nm = fmaxf(m, om);
    float a = (m == -INFINITY) ? 0.0f : expf(m - nm);
    float b = (om == -INFINITY) ? 0.0f : expf(om - nm);
    if (om != om || m != m) nm = NAN;             // NaN log-lik poisons the step like jnp.max does
    // (explicit mul / add, no fma: both partners of a butterfly must get the same bits from swapped arguments)
    s = __fadd_rn(__fmul_rn(s, a), __fmul_rn(os, b)); s2 = __fadd_rn(__fmul_rn(s2, __fmul_rn(a, a)), __fmul_rn(os2, __fmul_rn(b, b)));
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] = __fadd_rn(__fmul_rn(sx[k], a), __fmul_rn(osx[k], b));
    m = nm;
  }
  // ---- canonical reduction tree of the step statistics (pf_fused.cuh): the same binary tree over the particle index
  // whatever the grid, the warp assignment or the number of GPUs a filter is sharded over.
  __device__ __forceinline__ void merge_acc(const OnlineAcc& o) { merge(o.m, o.s, o.s2, o.sx); }
  __device__ __forceinline__ void load(const PfPartial* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    const float4 a = __ldcg(q);
    m = a.x; s = a.y; s2 = a.z;
    float v[8];
    const float4 b = __ldcg(q + 1);
    v[0] = b.x; v[1] = b.y; v[2] = b.z; v[3] = b.w;
    if (DX > 4) { const float4 c = __ldcg(q + 2); v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w; }
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] = v[k];
  }
  __device__ __forceinline__ void load_plain(const PfPartial* p) {     // shared or local memory
    m = p->m; s = p->s; s2 = p->s2;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] = p->sx[k];
  }
  __device__ __forceinline__ void store(PfPartial* p) const {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (k < DX) ? sx[k < DX ? k : 0] : 0.0f;
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = make_float4(m, s, s2, 0.0f);
    q[1] = make_float4(v[0], v[1], v[2], v[3]);
    if (DX > 4) q[2] = make_float4(v[4], v[5], v[6], v[7]);
  }
  // the 32 lane accumulators of one unit (512 consecutive particles, 16 per lane) -> the unit's partial, in every lane:
  // one common maximum, one rescale per lane, then plain butterfly sums (adjacent lanes first)
  __device__ __forceinline__ void unit_reduce() {
    const bool anynan = __any_sync(BF_FULL, m != m);
    const float mu = warp_max(m);
    const float f = (m == -INFINITY) ? 0.0f : mufu_ex2((m - mu) * 1.44269504088896340736f);
    s *= f; s2 *= f * f;
#pragma unroll
    for (int k = 0; k < DX; ++k) sx[k] *= f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      s = __fadd_rn(s, __shfl_xor_sync(BF_FULL, s, o)); s2 = __fadd_rn(s2, __shfl_xor_sync(BF_FULL, s2, o));
#pragma unroll
      for (int k = 0; k < DX; ++k) sx[k] = __fadd_rn(sx[k], __shfl_xor_sync(BF_FULL, sx[k], o));
    }
    m = anynan ? NAN : mu;
  }
  // butterfly of full merges, adjacent lanes first: the balanced binary tree over the lane index
  __device__ __forceinline__ void warp_tree_reduce() {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float om = __shfl_xor_sync(BF_FULL, m, o), os = __shfl_xor_sync(BF_FULL, s, o), os2 = __shfl_xor_sync(BF_FULL, s2, o);
      float osx[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) osx[k] = __shfl_xor_sync(BF_FULL, sx[k], o);
      merge(om, os, os2, osx);
    }
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

// normalised weight of particle i exactly as inference.py:1350-1353: (exp(ll - max) * w_prev) / sum
__device__ __forceinline__ float norm_weight(float ll, float wprev, float mx, float Z) {
  return __fdiv_rn(__fmul_rn(expf(__fsub_rn(ll, mx)), wprev), Z);
}
// same with the reciprocal of the sum hoisted (<= 1.5 ulp from the division; used consistently in every phase)
__device__ __forceinline__ float norm_weight_rcp(float ll, float wprev, float mx, float rZ) {
  return __fmul_rn(__fmul_rn(expf(__fsub_rn(ll, mx)), wprev), rZ);
}

#ifdef BF_TU_MAIN
// ---------------------------------------------------------------- k_scan_tiles (one CTA)
__global__ void __launch_bounds__(PF_TILE_WARPS * 32)
k_scan_tiles(const float* tilesum, float* Etile, int ntiles, const StepStats* cur) {
  extern __shared__ float s_scan[];
  if (cur && !cur->flag) return;
  cta_exclusive_scan<PF_TILE_WARPS>(tilesum, Etile, ntiles, s_scan);
}
#endif  // BF_TU_MAIN

// resampling position of output slot i (oracle/scan.py multinomial_positions / systematic_positions)
__device__ __forceinline__ float resample_position(int resampler, float total, float u, int i, int N) {
  if (resampler == 1) return __fmul_rn(total, __fdiv_rn(__fadd_rn((float)i, u), (float)N));
  return __fmul_rn(total, __fsub_rn(1.0f, u));
}

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

#ifdef BF_TU_MAIN
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
#endif  // BF_TU_MAIN

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
