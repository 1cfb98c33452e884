// pf_resident.cuh -- bootstrap particle filter with the whole filter state resident in shared memory.
//
// One CTA per independent Monte-Carlo trial runs ALL T steps of gaussfiltax/inference.py:1330-1366
// (propagate, weight, max-shift normalise, ESS test, resample) without leaving the SM: particles
// (ping-pong), weights, the canonical cdf and the ancestor indices live in shared memory, the only
// HBM traffic is the supplied randomness (if any) and the requested per-step outputs.  This is the
// path for the reference's own problem sizes (N = 100 .. a few thousand, many trials: the
// `for i in range(Nsim)` loops of docs/experiments/BOT_Experiment_script.py:89 become the grid).
#pragma once
#include "pf_kernels.cuh"

struct PfResidentArgs {
  int N, T, t_offset, resume, resampler, emissions_per_trial;
  int trial_offset;                  // global index of trial 0 (Philox counters)
  float ess_threshold;
  unsigned long long seed;
  const float* emissions; const float* inputs;
  const float* init_particles; const float* init_weights;
  const float* noise; const float* res_u;
  float* carry_x; float* carry_w; int* carry_uniform;   // handle-owned carry [B,N,DX], [B,N], [B]
  float* out_particles; float* out_weights; float* out_propagated; float* out_norm_weights; int* out_ancestors;
  float* out_mean; float* out_loglik; float* out_ess; int* out_resampled;
  float* final_particles; float* final_weights;
};

template <int NWARPS, int K>
__device__ __forceinline__ void block_sum_k(float (&v)[K], float* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) s_red[warp * K + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float y = (lane < NWARPS) ? s_red[lane * K + k] : 0.0f;
    v[k] = warp_sum(y);
  }
}
template <int NWARPS>
__device__ __forceinline__ float block_max(float v, float* s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float y = (lane < NWARPS) ? s_red[lane] : -INFINITY;
  return warp_max(y);
}

static inline size_t pf_resident_smem_bytes(int N, int DX, int nwarps) {
  const int tile = nwarps * 128;
  const int nchunk = (N + 1 + tile - 1) / tile;
  size_t fl = (size_t)2 * N * DX + N + (N + 1) + N /*anc*/ + (nchunk + 1 + nwarps) + 32 * (2 + BF_MAXD) + 8;
  return fl * sizeof(float) + 64;
}

template <class M, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
k_pf_resident(const __grid_constant__ ModelConst mc, const __grid_constant__ PfResidentArgs a) {
  constexpr int DX = M::DX, DQ = M::DQ, DY = M::DY;
  constexpr int NT = NWARPS * 32;
  extern __shared__ __align__(16) float smem[];
  const int N = a.N, T = a.T, b = blockIdx.x, tid = threadIdx.x;
  const int Npad = (N + 3) & ~3;
  float* Xa = smem;                         // current particles [N,DX]
  float* Xb = Xa + (size_t)Npad * DX;       // propagated particles
  float* Wb = Xb + (size_t)Npad * DX;       // weights
  float* Eb = Wb + Npad;                    // log-liks, then exclusive cdf [N+1]
  int* ANC = reinterpret_cast<int*>(Eb + Npad + 4);
  float* s_red = reinterpret_cast<float*>(ANC + Npad);     // 32*(1+DX) reduction scratch
  float* s_scan = s_red + 32 * (2 + BF_MAXD);              // scan scratch
  const float invN = 1.0f / (float)N;
  const int levels = bisect_levels(N);

  // ---- load the carry ----
  bool w_uni;
  {
    const float* x0 = a.resume ? a.carry_x + (size_t)b * N * DX : (a.init_particles ? a.init_particles + (size_t)b * N * DX : nullptr);
    const float* w0 = a.resume ? a.carry_w + (size_t)b * N : (a.init_weights ? a.init_weights + (size_t)b * N : nullptr);
    w_uni = a.resume ? (a.carry_uniform[b] != 0) : (a.init_weights == nullptr);
    for (int i = tid; i < N; i += NT) {
      if (x0) {
#pragma unroll
        for (int k = 0; k < DX; ++k) Xa[i * DX + k] = x0[(size_t)i * DX + k];
      } else {
        float z[DX];
        philox_normals<DX>(a.seed, BF_STREAM_INIT, 0u, (uint32_t)(a.trial_offset + b), (uint64_t)i, z);
#pragma unroll
        for (int r = 0; r < DX; ++r) {
          float s = mc.m0[r];
#pragma unroll
          for (int c = 0; c <= r; ++c) s += mc.LP0[r * DX + c] * z[c];
          Xa[i * DX + r] = s;
        }
      }
      Wb[i] = (w0 && !w_uni) ? w0[i] : invN;
    }
  }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const float* yp = a.emissions + ((size_t)(a.emissions_per_trial ? b : 0) * T + t) * DY;
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = yp[k];
    const float u = a.inputs ? a.inputs[t] : 0.0f;
    const size_t bt = (size_t)b * T + t;

    // 1. propagate + log-lik                                     (inference.py:1342-1349)
    float tmax = -INFINITY;
    for (int i = tid; i < N; i += NT) {
      float x[DX], q[DQ], xt[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) x[k] = Xa[i * DX + k];
      if (a.noise) {
        const float* np_ = a.noise + (((size_t)t * gridDim.x + b) * N + i) * DQ;
#pragma unroll
        for (int k = 0; k < DQ; ++k) q[k] = np_[k];
      } else {
        float z[DQ];
        philox_normals<DQ>(a.seed, BF_STREAM_NOISE, (uint32_t)(t + a.t_offset), (uint32_t)(a.trial_offset + b), (uint64_t)i, z);
#pragma unroll
        for (int r = 0; r < DQ; ++r) {
          float s = mc.q0[r];
#pragma unroll
          for (int c = 0; c <= r; ++c) s += mc.LQ[r * DQ + c] * z[c];
          q[r] = s;
        }
      }
      M::f(x, q, u, mc, xt);
#pragma unroll
      for (int k = 0; k < DX; ++k) Xb[i * DX + k] = xt[k];
      if (a.out_propagated) {
#pragma unroll
        for (int k = 0; k < DX; ++k) a.out_propagated[(bt * N + i) * DX + k] = xt[k];
      }
      const float ll = emission_log_prob<M>(xt, y, u, mc);
      Eb[i] = ll;
      tmax = fmaxf(tmax, ll);
    }
    const float mx = block_max<NWARPS>(tmax, s_red);

    // 2. lls -= max ; w = exp(lls) * w                           (inference.py:1350-1352)
    float zs[1] = {0.0f};
    for (int i = tid; i < N; i += NT) {
      const float e = expf(Eb[i] - mx) * (w_uni ? invN : Wb[i]);
      Wb[i] = e;
      zs[0] += e;
    }
    block_sum_k<NWARPS, 1>(zs, s_red);
    const float Z = zs[0];

    // 3. w /= sum(w) ; ESS ; weighted mean                       (inference.py:1353, 1356)
    float acc[1 + DX];
#pragma unroll
    for (int k = 0; k < 1 + DX; ++k) acc[k] = 0.0f;
    for (int i = tid; i < N; i += NT) {
      const float w = Wb[i] / Z;
      Wb[i] = w;
      acc[0] += w * w;
#pragma unroll
      for (int k = 0; k < DX; ++k) acc[1 + k] += w * Xb[i * DX + k];
      if (a.out_norm_weights) a.out_norm_weights[bt * N + i] = w;
    }
    block_sum_k<NWARPS, 1 + DX>(acc, s_red);
    const float ess = 1.0f / acc[0];
    const bool flag = ess < a.ess_threshold * (float)N;
    if (tid == 0) {
      if (a.out_loglik) a.out_loglik[bt] = mx + logf(Z);
      if (a.out_ess) a.out_ess[bt] = ess;
      if (a.out_resampled) a.out_resampled[bt] = flag ? 1 : 0;
      if (a.out_mean) {
#pragma unroll
        for (int k = 0; k < DX; ++k) a.out_mean[bt * DX + k] = acc[1 + k];
      }
    }
    __syncthreads();

    // 4. lax.cond(resample_cond, _resample, identity)            (inference.py:1357, utils.py:207-214)
    if (flag) {
      cta_exclusive_scan<NWARPS>(Wb, Eb, N, s_scan);
      const float total = Eb[N];
      float u0 = 0.0f;
      if (a.resampler == 1)
        u0 = a.res_u ? a.res_u[(size_t)t * gridDim.x + b]
                     : philox_uniform(a.seed, BF_STREAM_RESAMPLE, (uint32_t)(t + a.t_offset), (uint32_t)(a.trial_offset + b), 0ull);
      for (int i = tid; i < N; i += NT) {
        float uu = u0;
        if (a.resampler == 0)
          uu = a.res_u ? a.res_u[((size_t)t * gridDim.x + b) * N + i]
                       : philox_uniform(a.seed, BF_STREAM_RESAMPLE, (uint32_t)(t + a.t_offset), (uint32_t)(a.trial_offset + b), (uint64_t)i);
        const float r = resample_position(a.resampler, total, uu, i, N);
        ANC[i] = bisect_left_exclusive(Eb, N, levels, r);
      }
      __syncthreads();
      for (int i = tid; i < N; i += NT) {
        const int src = ANC[i];
#pragma unroll
        for (int k = 0; k < DX; ++k) Xa[i * DX + k] = Xb[src * DX + k];
        if (a.out_ancestors) a.out_ancestors[bt * N + i] = src;
      }
      w_uni = true;
    } else {
      float* tmp = Xa; Xa = Xb; Xb = tmp;
      w_uni = false;
      if (a.out_ancestors)
        for (int i = tid; i < N; i += NT) a.out_ancestors[bt * N + i] = i;
    }
    __syncthreads();

    // 5. outputs of the step (post-resample)                     (inference.py:1359-1362)
    if (a.out_particles) {
      float* o = a.out_particles + bt * N * DX;
      for (int j = tid; j < N * DX; j += NT) o[j] = Xa[j];
    }
    if (a.out_weights) {
      float* o = a.out_weights + bt * N;
      for (int i = tid; i < N; i += NT) o[i] = w_uni ? invN : Wb[i];
    }
  }

  // ---- store the carry ----
  for (int j = tid; j < N * DX; j += NT) {
    a.carry_x[(size_t)b * N * DX + j] = Xa[j];
    if (a.final_particles) a.final_particles[(size_t)b * N * DX + j] = Xa[j];
  }
  for (int i = tid; i < N; i += NT) {
    const float w = w_uni ? invN : Wb[i];
    a.carry_w[(size_t)b * N + i] = w;
    if (a.final_weights) a.final_weights[(size_t)b * N + i] = w;
  }
  if (tid == 0) a.carry_uniform[b] = w_uni ? 1 : 0;
}
