// pf_fused.cuh -- bootstrap particle filter, large-N path: ONE persistent cooperative kernel runs all T
// steps of gaussfiltax/inference.py:1330-1366 with the particle state streamed through HBM (sm_100a).
//
// Grid = every co-resident CTA of the device (cooperative launch), tiles of NWARPS*128 consecutive particles
// are owned by the same CTA (and the same 4 particles by the same thread) in every phase, so a step is
//
//   A  ancestor of slot i from the PREVIOUS step's cdf (the resampling gather is fused into the propagate:
//      x = X[t-1][a_i], never materialised), x~ = f(x, q, u), ll = log p(y | x~), CTA partial of the online
//      (max, sum e, sum e^2, sum e x~) accumulator                                           -- grid sync --
//   B  every CTA folds the partials in the same fixed order -> max, Z, log-lik increment, ESS, resample
//      flag, weighted mean (bit-identical in every CTA: the branch below is grid-uniform);
//      w = exp(ll - max) w_prev / Z ; no resampling: store w and go to the next step WITHOUT a grid sync;
//      resampling: canonical tree sum of each tile of w                                      -- grid sync --
//   C1 CTA 0: canonical exclusive scan of the tile sums                                      -- grid sync --
//   C2 w again (recomputed from ll, bit-identical), canonical exclusive scan inside each tile -> E[0..N]
//      (cdf[j] = E[j+1], lax.associative_scan order: oracle/scan.py)                         -- grid sync --
//
// and after the last step one more gather materialises the carry.  HBM traffic of a resampling step with
// in-kernel Philox noise, per particle: A read x 4DX + write x~ 4DX + write ll 4 (+ cdf probes, cached);
// B read ll 4; C2 read ll 4 + write E 4  ->  8 DX + 16 bytes (48 B for DX = 4).
//
// The ancestor search is jnp.searchsorted's fixed bisection (oracle/scan.py searchsorted_bisect), the probe
// sequence is part of the contract because the tree-ordered fp32 cdf is not monotone in its last bit.
#pragma once
#include <cooperative_groups.h>
#include "pf_kernels.cuh"

namespace cg = cooperative_groups;

#define PF_FUSED_WARPS 16                   // 512 threads, 2 CTAs per SM, tiles of 2048 particles

struct PfFusedArgs {
  int N, T, t_offset, trial, B;
  int resume, w_uniform0, resampler;
  int ntiles, tiles_per_cta, log2n;          // log2n >= 0 when N is a power of two (fast bisection), else -1
  float ess_threshold;
  unsigned long long seed;
  const float* emissions;                    // [T,DY] of this trial
  const float* inputs;                       // [T] or NULL
  const float* x_init;                       // [N,DX] particles entering step 0
  float* xbuf0; float* xbuf1;                // [N,DX] ping-pong
  float* W;                                  // [N] carried weights (valid when the carry is not uniform)
  float* LL;                                 // [ntiles*TILE]
  float* E;                                  // [ntiles*TILE] exclusive canonical cdf, E[N] = total
  float* tilesum;                            // [ntiles]
  float* Etile;                              // [ntiles+1]
  PfPartial* partials;                       // [2*gridDim.x]
  StepStats* carry_stats;                    // [1] in (resume) / out
  const float* noise;                        // + (t*B + trial)*N*DQ, or NULL (Philox)
  const float* res_u;                        // multinomial + (t*B + trial)*N ; systematic + t*B + trial ; or NULL
  float* out_particles; float* out_weights; float* out_propagated; float* out_norm_weights; int* out_ancestors;
  float* out_mean; float* out_loglik; float* out_ess; int* out_resampled;      // already offset to (trial, t=0)
  float* carry_x; float* final_particles; float* final_weights;               // carry_w is W
  unsigned long long* phase_ns;             // [5] accumulated wall time of phases A, B, C1, C2, final as seen by CTA 0
};

// DQ standard normals for each of the 4 particles of quad `quad` (= particle index / 4): call k supplies
// component k of all four.  Counter = (quad, trial, step, stream | k << 8).
template <int DQ>
__device__ __forceinline__ void philox_quad_normals(unsigned long long seed, uint32_t stream, uint32_t t, uint32_t b,
                                                    uint32_t quad, float z[4][DQ]) {
#pragma unroll
  for (int k = 0; k < DQ; ++k) {
    const uint4x r = philox4x32(quad, b, t, stream | ((uint32_t)k << 8), (uint32_t)seed, (uint32_t)(seed >> 32));
    box_muller(r.x, r.y, z[0][k], z[1][k]);
    box_muller(r.z, r.w, z[2][k], z[3][k]);
  }
}
__device__ __forceinline__ void philox_quad_uniforms(unsigned long long seed, uint32_t stream, uint32_t t, uint32_t b,
                                                     uint32_t quad, float u[4]) {
  const uint4x r = philox4x32(quad, b, t, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  u[0] = u01(r.x); u[1] = u01(r.y); u[2] = u01(r.z); u[3] = u01(r.w);
}

// The cdf is written by this same kernel (phase C2) and read here after a grid sync: plain coherent loads
// (no __restrict__/ld.global.nc), L1-cached -- cg's grid.sync() fences and invalidates L1.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int fused_bisect_left_exclusive(const float* ex, int n, int n_levels, float r) {
  int low = 0, high = n;
  for (int l = 0; l < n_levels; ++l) {
    const int mid = (low + high) >> 1;
    const int m = mid < n - 1 ? mid : n - 1;
    const bool go_left = r <= ex[m + 1];
    high = go_left ? mid : high;
    low = go_left ? low : mid;
  }
  return high < n - 1 ? high : n - 1;
}

// Fixed bisection for n = 2^log2n, four queries at once.  The widths halve exactly, so only `low` has to be
// tracked; same probes, same comparisons, same result as bisect_left_exclusive (scan.cuh) for power-of-two n.
// The four dependent-load chains are interleaved by hand (ptxas keeps them serial otherwise).
__device__ __forceinline__ void bisect4_left_exclusive_pow2(const float* ex, int log2n, const float r[4], int anc[4]) {
  const float* ex1 = ex + 1;
  unsigned low0 = 0, low1 = 0, low2 = 0, low3 = 0;
  for (unsigned s = 1u << (log2n - 1); s != 0; s >>= 1) {
    const float v0 = ex1[low0 + s], v1 = ex1[low1 + s], v2 = ex1[low2 + s], v3 = ex1[low3 + s];
    if (!(r[0] <= v0)) low0 += s;
    if (!(r[1] <= v1)) low1 += s;
    if (!(r[2] <= v2)) low2 += s;
    if (!(r[3] <= v3)) low3 += s;
  }
  const float v0 = ex1[low0], v1 = ex1[low1], v2 = ex1[low2], v3 = ex1[low3];
  const unsigned nm1 = (1u << log2n) - 1u;
  anc[0] = (int)min(low0 + ((r[0] <= v0) ? 0u : 1u), nm1);
  anc[1] = (int)min(low1 + ((r[1] <= v1) ? 0u : 1u), nm1);
  anc[2] = (int)min(low2 + ((r[2] <= v2) ? 0u : 1u), nm1);
  anc[3] = (int)min(low3 + ((r[3] <= v3) ? 0u : 1u), nm1);
}

template <int DX>
__device__ __forceinline__ void load_state_cg(const float* base, size_t i, float* x) {
  if constexpr (DX % 4 == 0) {
    const float4* p = reinterpret_cast<const float4*>(base) + i * (DX / 4);
#pragma unroll
    for (int k = 0; k < DX / 4; ++k) { float4 v = __ldcg(p + k); x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w; }
  } else if constexpr (DX % 2 == 0) {
    const float2* p = reinterpret_cast<const float2*>(base) + i * (DX / 2);
#pragma unroll
    for (int k = 0; k < DX / 2; ++k) { float2 v = __ldcg(p + k); x[2 * k] = v.x; x[2 * k + 1] = v.y; }
  } else {
#pragma unroll
    for (int k = 0; k < DX; ++k) x[k] = __ldcg(base + i * DX + k);
  }
}

template <class M, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 2048 / (NWARPS * 32) > 2 ? 2 : 2048 / (NWARPS * 32))
k_pf_fused(const __grid_constant__ ModelConst mc, const __grid_constant__ PfFusedArgs a) {
  constexpr int DX = M::DX, DQ = M::DQ, DY = M::DY;
  constexpr int NT = NWARPS * 32, TILE = NWARPS * 128;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float s_dyn[];                    // scan scratch of CTA 0 (phase C1)
  __shared__ float s_warp[2][NWARPS];
  __shared__ PfPartial s_part[NWARPS];
  __shared__ StepStats s_stats;

  const int N = a.N, T = a.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int tile0 = blockIdx.x * a.tiles_per_cta;
  const int tile1 = min(tile0 + a.tiles_per_cta, a.ntiles);
  const float invN = 1.0f / (float)N;
  const int levels = bisect_levels(N);

  bool w_uni = a.resume ? (a.carry_stats->w_uniform != 0) : (a.w_uniform0 != 0);
  bool gather = false;                                // the carry entering the call is materialised
  const float* x_prev = a.x_init;
  int sw = 0;                                         // s_warp double-buffer index

  // ancestors of the 4 slots base..base+3 for the resampling done at (local) step ts
  auto ancestors4 = [&](int base, int ts, int anc[4]) {
    const float total = a.E[N];
    float u[4];
    if (a.resampler == 1) {
      const float u0 = a.res_u ? a.res_u[(size_t)ts * a.B]
                               : philox_uniform(a.seed, BF_STREAM_RESAMPLE, (uint32_t)(ts + a.t_offset), (uint32_t)a.trial, 0ull);
      u[0] = u[1] = u[2] = u[3] = u0;
    } else if (a.res_u) {
      const float* p = a.res_u + (size_t)ts * a.B * N;
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = (base + k < N) ? p[base + k] : 0.0f;
    } else {
      philox_quad_uniforms(a.seed, BF_STREAM_RESAMPLE, (uint32_t)(ts + a.t_offset), (uint32_t)a.trial, (uint32_t)(base >> 2), u);
    }
    float r[4];
    if (a.log2n >= 0) {
      // (i + u) / 2^k is an exact scaling: multiplying by 2^-k gives the same bits as the oracle's division
      const float invn = __int_as_float((127 - a.log2n) << 23);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        r[k] = (a.resampler == 1) ? __fmul_rn(total, __fmul_rn(__fadd_rn((float)(base + k), u[k]), invn))
                                  : __fmul_rn(total, __fsub_rn(1.0f, u[k]));
      bisect4_left_exclusive_pow2(a.E, a.log2n, r, anc);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] = resample_position(a.resampler, total, u[k], base + k, N);
#pragma unroll
      for (int k = 0; k < 4; ++k) anc[k] = fused_bisect_left_exclusive(a.E, N, levels, r[k]);
    }
  };

  const bool timer = (blockIdx.x == 0 && tid == 0 && a.phase_ns != nullptr);
  unsigned long long tmark = timer ? globaltimer_ns() : 0ull;
  auto lap = [&](int phase) {
    if (timer) { const unsigned long long now = globaltimer_ns(); a.phase_ns[phase] += now - tmark; tmark = now; }
  };

  for (int t = 0; t < T; ++t) {
    float* x_cur = (t & 1) ? a.xbuf1 : a.xbuf0;
    const float u_in = a.inputs ? a.inputs[t] : 0.0f;
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = a.emissions[(size_t)t * DY + k];

    // ------------------------------------------------------------------ phase A
    OnlineAcc<DX> acc; acc.init();
    for (int tile = tile0; tile < tile1; ++tile) {
      const int base = tile * TILE + tid * 4;
      if (base >= N) continue;                         // no barrier inside this loop
      const int nv = min(4, N - base);
      int anc[4] = {base, base + 1, base + 2, base + 3};
      if (gather) ancestors4(base, t - 1, anc);
      if (t > 0 && a.out_ancestors) {
        int* o = a.out_ancestors + (size_t)(t - 1) * N + base;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < nv) o[k] = anc[k];
      }
      float z[4][DQ];
      if (!a.noise) philox_quad_normals<DQ>(a.seed, BF_STREAM_NOISE, (uint32_t)(t + a.t_offset), (uint32_t)a.trial, (uint32_t)(base >> 2), z);
      float llv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nv) {
          const int i = base + k;
          float x[DX], q[DQ], xt[DX];
          load_state_cg<DX>(x_prev, (size_t)anc[k], x);
          if (t > 0 && a.out_particles) store_state<DX>(a.out_particles + (size_t)(t - 1) * N * DX, (size_t)i, x);
          if (a.noise) {
            const float* np_ = a.noise + (size_t)t * a.B * N * DQ + (size_t)i * DQ;
#pragma unroll
            for (int c = 0; c < DQ; ++c) q[c] = np_[c];
          } else {
#pragma unroll
            for (int r = 0; r < DQ; ++r) {             // q = q0 + L_Q z  (models.py:83, TFP MVN sample)
              float s = mc.q0[r];
#pragma unroll
              for (int c = 0; c <= r; ++c) s += mc.LQ[r * DQ + c] * z[k][c];
              q[r] = s;
            }
          }
          M::f(x, q, u_in, mc, xt);
          store_state<DX>(x_cur, (size_t)i, xt);
          if (a.out_propagated) store_state<DX>(a.out_propagated + (size_t)t * N * DX, (size_t)i, xt);
          const float ll = emission_log_prob<M>(xt, y, u_in, mc);
          llv[k] = ll;
          acc.add(ll, w_uni ? invN : a.W[i], xt);
        }
      }
      *reinterpret_cast<float4*>(a.LL + base) = make_float4(llv[0], llv[1], llv[2], llv[3]);
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
      if (lane < NWARPS) acc.merge(s_part[lane].m, s_part[lane].s, s_part[lane].s2, s_part[lane].sx);
      acc.warp_reduce();
      if (lane == 0) {
        PfPartial* p = &a.partials[(size_t)(t & 1) * G + blockIdx.x];
        p->m = acc.m; p->s = acc.s; p->s2 = acc.s2;
#pragma unroll
        for (int k = 0; k < DX; ++k) p->sx[k] = acc.sx[k];
      }
    }
    grid.sync();
    lap(0);

    // ------------------------------------------------------------------ phase B
    acc.init();
    for (int b = tid; b < G; b += NT) {               // same order in every CTA: bit-identical stats
      const PfPartial* p = &a.partials[(size_t)(t & 1) * G + b];
      float sx[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) sx[k] = __ldcg(&p->sx[k]);
      acc.merge(__ldcg(&p->m), __ldcg(&p->s), __ldcg(&p->s2), sx);
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
      if (lane < NWARPS) acc.merge(s_part[lane].m, s_part[lane].s, s_part[lane].s2, s_part[lane].sx);
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
        s_stats = st;
        if (blockIdx.x == 0) {
          if (a.out_loglik) a.out_loglik[t] = st.loglik;
          if (a.out_ess) a.out_ess[t] = st.ess;
          if (a.out_resampled) a.out_resampled[t] = st.flag;
          if (a.out_mean) {
#pragma unroll
            for (int k = 0; k < DX; ++k) a.out_mean[(size_t)t * DX + k] = st.mean[k];
          }
        }
      }
    }
    __syncthreads();
    const float mx = s_stats.mx, rZ = 1.0f / s_stats.Z;
    const bool flag = s_stats.flag != 0;

    // normalised weights of this thread's 4 leaves in `tile`, exactly as inference.py:1350-1353
    auto weights4 = [&](int tile, float v[4]) {
      const int base = tile * TILE + tid * 4;
      const float4 l4 = *reinterpret_cast<const float4*>(a.LL + base);
      const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float wp = (base + k < N) ? (w_uni ? invN : a.W[base + k]) : 0.0f;
        v[k] = (base + k < N) ? norm_weight_rcp(ll[k], wp, mx, rZ) : 0.0f;
      }
    };

    for (int tile = tile0; tile < tile1; ++tile) {
      const int base = tile * TILE + tid * 4;
      float v[4];
      weights4(tile, v);
      if (a.out_norm_weights) {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (base + k < N) a.out_norm_weights[(size_t)t * N + base + k] = v[k];
      }
      if (!flag) {                       // identity branch of lax.cond (inference.py:1357): carry the weights
#pragma unroll
        for (int k = 0; k < 4; ++k) if (base + k < N) {
          a.W[base + k] = v[k];
          if (a.out_weights) a.out_weights[(size_t)t * N + base + k] = v[k];
        }
      } else {
        if (a.out_weights) {             // utils.py:212: weights = 1/N after resampling
#pragma unroll
          for (int k = 0; k < 4; ++k) if (base + k < N) a.out_weights[(size_t)t * N + base + k] = invN;
        }
        const float ts = tile_tree_sum<NWARPS>(v, s_warp[sw]);
        sw ^= 1;
        if (tid == 0) a.tilesum[tile] = ts;
      }
    }
    x_prev = x_cur;
    if (!flag) { w_uni = false; gather = false; lap(1); continue; }     // grid-uniform branch
    grid.sync();
    lap(1);

    // ------------------------------------------------------------------ phase C1 (CTA 0)
    if (blockIdx.x == 0) cta_exclusive_scan<NWARPS>(a.tilesum, a.Etile, a.ntiles, s_dyn);
    grid.sync();
    lap(2);

    // ------------------------------------------------------------------ phase C2
    for (int tile = tile0; tile < tile1; ++tile) {
      float v[4], e[4];
      weights4(tile, v);
      tile_exclusive_scan<NWARPS>(v, __ldcg(&a.Etile[tile]), e, s_warp[sw]);
      sw ^= 1;
      *reinterpret_cast<float4*>(a.E + (size_t)tile * TILE + tid * 4) = make_float4(e[0], e[1], e[2], e[3]);
    }
    grid.sync();
    lap(3);
    w_uni = true; gather = true;
  }

  // ------------------------------------------------------------------ materialise the carry
  for (int tile = tile0; tile < tile1; ++tile) {
    const int base = tile * TILE + tid * 4;
    if (base >= N) continue;
    const int nv = min(4, N - base);
    int anc[4] = {base, base + 1, base + 2, base + 3};
    if (gather) ancestors4(base, T - 1, anc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nv) {
        const int i = base + k;
        float x[DX];
        load_state_cg<DX>(x_prev, (size_t)anc[k], x);
        store_state<DX>(a.carry_x, (size_t)i, x);
        if (a.final_particles) store_state<DX>(a.final_particles, (size_t)i, x);
        if (a.out_particles) store_state<DX>(a.out_particles + (size_t)(T - 1) * N * DX, (size_t)i, x);
        if (a.out_ancestors) a.out_ancestors[(size_t)(T - 1) * N + i] = anc[k];
        const float w = w_uni ? invN : a.W[i];
        if (w_uni) a.W[i] = invN;
        if (a.final_weights) a.final_weights[i] = w;
      }
    }
  }
  lap(4);
  if (blockIdx.x == 0 && tid == 0) {
    StepStats st = s_stats;
    st.flag = 0; st.w_uniform = w_uni ? 1 : 0;
    *a.carry_stats = st;
  }
}
