// gsf_kernels.cuh -- Gaussian-sum (mixture of EKFs) filter, persistent over time (sm_100a).
//
// Restates gaussfiltax/inference.py:333-365 (`_step` of gaussian_sum_filter): per component
// `_condition_on` (:72-105) on the carried prediction, the max-shifted evidence reweighting across the
// M components (:347-350), then `_predict` (:51-70).  One CTA per independent trial, one thread per
// mixture component; a component's weight, mean and covariance stay in that thread's registers for all
// T steps, so the only per-step HBM traffic is the requested outputs.  The only cross-component coupling
// (one max and one sum per step) is a warp-shuffle + shared-memory CTA reduction.
//
// Arithmetic mirrors the reference, including its quirks:
//   K  = solve(S + 1e-6 [every element], H P)^T          utils.py:256-259 psd_solve, LU with partial pivoting
//   P+ = P - (K S) K^T with the UN-jittered S            inference.py:101
//   ll = MVN(h(m), S).log_prob(y), Cholesky of the un-jittered S   inference.py:104 (TFP)
//   no re-symmetrisation of P+.
// Emission noise is additive (H_r = I) for every registered model except MSV, whose H_r(x, u) is diagonal (models.cuh);
// F_q is constant for all of them, so F_q Q F_q^T is computed once per CTA.
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "../../include/bfilt.h"

struct GsfArgs {
  int M, T, resume, emissions_per_trial, flags;
  const float* emissions; const float* inputs;
  const float* init_means; const float* init_covs; const float* init_weights;
  float* carry_w; float* carry_m; float* carry_P;       // handle-owned carry [B,M], [B,M,DX], [B,M,DX,DX]
  float* out_weights; float* out_means; float* out_covs; float* out_pred_means; float* out_pred_covs;
  float* out_loglik; float* out_mix_mean;
  const float* Q_t; const float* R_t;                   // [T,DQ,DQ] / [T,DY,DY] time-varying covariances or NULL (k_gsf only)
  const float* mc_draws;                                // legacy MCF: [B,T,M,2,np,DX] offsets x - m, or NULL (k_ugsf mode 2)
};

// C[RxC] = A[RxK] @ B[KxC]
template <int R, int K, int C>
__device__ __forceinline__ void mm(const float* A, const float* B, float* Cm) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[i * K + k] * B[k * C + j];
      Cm[i * C + j] = s;
    }
}
// C[RxC] = A[RxK] @ B[CxK]^T
template <int R, int K, int C>
__device__ __forceinline__ void mmt(const float* A, const float* B, float* Cm) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k];
      Cm[i * C + j] = s;
    }
}

// Solve A X = B in place (B <- X) by Gaussian elimination with partial pivoting, A [NxN], B [NxC].
// Row swaps are predicated so every index stays compile-time (registers, no local memory).
template <int N, int C>
__device__ __forceinline__ void lu_solve(float* A, float* B) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int p = k;
    float best = fabsf(A[k * N + k]);
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      float v = fabsf(A[r * N + k]);
      if (v > best) { best = v; p = r; }
    }
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      const bool sw = (p == r);
#pragma unroll
      for (int c = 0; c < N; ++c) { float a0 = A[k * N + c], a1 = A[r * N + c]; A[k * N + c] = sw ? a1 : a0; A[r * N + c] = sw ? a0 : a1; }
#pragma unroll
      for (int c = 0; c < C; ++c) { float b0 = B[k * C + c], b1 = B[r * C + c]; B[k * C + c] = sw ? b1 : b0; B[r * C + c] = sw ? b0 : b1; }
    }
    const float piv = A[k * N + k];
#pragma unroll
    for (int r = k + 1; r < N; ++r) {
      const float l = A[r * N + k] / piv;
#pragma unroll
      for (int c = k + 1; c < N; ++c) A[r * N + c] -= l * A[k * N + c];
#pragma unroll
      for (int c = 0; c < C; ++c) B[r * C + c] -= l * B[k * C + c];
    }
  }
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float s = B[k * C + c];
#pragma unroll
      for (int j = k + 1; j < N; ++j) s -= A[k * N + j] * B[j * C + c];
      B[k * C + c] = s / A[k * N + k];
    }
  }
}

// log N(y; yhat, S) with TFP's formula (Cholesky, forward substitution)
template <int N>
__device__ __forceinline__ float mvn_logpdf_chol(const float* S, const float* d) {
  float L[N * N], iL[N];
  float logdet = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float s = S[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i * N + k] * L[j * N + k];
      // pivot: L_ii = sqrt(s), 1/L_ii = rsqrt(s) (2 ulp), log L_ii = log(s)/2: no division, one log
      if (i == j) { iL[i] = rsqrtf(s); L[i * N + i] = s * iL[i]; logdet += 0.5f * logf(s); }
      else L[i * N + j] = s * iL[j];
    }
  }
  float ss = 0.0f, z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = d[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s -= L[i * N + j] * z[j];
    z[i] = s * iL[i];
    ss += z[i] * z[i];
  }
  return -0.5f * ss - logdet - 0.5f * (float)N * 1.8378770664093453f;
}

// one component: `_condition_on` (inference.py:72-105).  m,P in/out; returns ll.
template <class Mo>
__device__ __forceinline__ float ekf_condition(float* m, float* P, const float* y, float u, const ModelConst& mc, float jitter,
                                               const float* Rm) {
  constexpr int DX = Mo::DX, DY = Mo::DY;
  float H[DY * DX], HP[DY * DX], S[DY * DY], Sj[DY * DY], X[DY * DX];
  Mo::Hx(m, mc.r0, u, mc, H);
  mm<DY, DX, DX>(H, P, HP);                         // H_x @ P
  mmt<DY, DX, DY>(HP, H, S);                        // (H_x P) H_x^T
  float hr[DY];
#pragma unroll
  for (int i = 0; i < DY; ++i) hr[i] = 1.0f;
  if constexpr (nonadditive_r<Mo>::value) Mo::Hr_diag(m, u, mc, hr);   // H_r = dh/dr at (m, r0, u): diagonal in the registry
#pragma unroll
  for (int i = 0; i < DY; ++i)
#pragma unroll
    for (int j = 0; j < DY; ++j) {
      const float rr = nonadditive_r<Mo>::value ? hr[i] * Rm[i * DY + j] * hr[j] : Rm[i * DY + j];
      S[i * DY + j] = rr + S[i * DY + j];                    // H_r R H_r^T + H_x P H_x^T  (:99)
      Sj[i * DY + j] = S[i * DY + j] + jitter;               // psd_solve: + 1e-6 on EVERY element
    }
#pragma unroll
  for (int i = 0; i < DY * DX; ++i) X[i] = HP[i];
  lu_solve<DY, DX>(Sj, X);                          // X = solve(S', H P);  K = X^T   (:100)
  float KS[DX * DY];                                // K @ S
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DY; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < DY; ++k) s += X[k * DX + i] * S[k * DY + j];
      KS[i * DY + j] = s;
    }
#pragma unroll
  for (int i = 0; i < DX; ++i)                      // P - (K S) K^T           (:101)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < DY; ++k) s += KS[i * DY + k] * X[k * DX + j];
      P[i * DX + j] -= s;
    }
  float yhat[DY], d[DY];
  Mo::h(m, mc.r0, u, mc, yhat);
#pragma unroll
  for (int i = 0; i < DY; ++i) d[i] = y[i] - yhat[i];
#pragma unroll
  for (int i = 0; i < DX; ++i) {                    // m + K (y - h(m))        (:102)
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < DY; ++k) s += X[k * DX + i] * d[k];
    m[i] += s;
  }
  return mvn_logpdf_chol<DY>(S, d);                 // (:103-104)
}

// one component: `_predict` (inference.py:51-70).  GQG = F_q Q F_q^T (constant F_q).
// cov_plus_p: the legacy gausssumfilt.GaussSumFilt adds the covariance itself instead of Q (gausssumfilt.py:59)
template <class Mo>
__device__ __forceinline__ void ekf_predict(float* m, float* P, float u, const ModelConst& mc, const float* GQG, bool cov_plus_p = false) {
  constexpr int DX = Mo::DX;
  float F[DX * DX], FP[DX * DX], mn[DX];
  Mo::Fx(m, mc.q0, u, mc, F);
  Mo::f(m, mc.q0, u, mc, mn);
  mm<DX, DX, DX>(F, P, FP);
#pragma unroll
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < DX; ++k) s += FP[i * DX + k] * F[j * DX + k];
      P[i * DX + j] = s + (cov_plus_p ? P[i * DX + j] : GQG[i * DX + j]);
    }
#pragma unroll
  for (int i = 0; i < DX; ++i) m[i] = mn[i];
}

template <int K>
__device__ __forceinline__ void store_vec(float* dst, const float* v) {
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int k = 0; k < K / 4; ++k) reinterpret_cast<float4*>(dst)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int k = 0; k < K / 2; ++k) reinterpret_cast<float2*>(dst)[k] = make_float2(v[2 * k], v[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) dst[k] = v[k];
  }
}

// TV: time-varying noise covariances (a.Q_t / a.R_t).  A separate instantiation: with the per-step reload in the one
// kernel the constant case ran 18 % slower (BASELINE configs[1]).
template <class Mo, bool TV>
__global__ void k_gsf(const __grid_constant__ ModelConst mc, const __grid_constant__ GsfArgs a) {
  constexpr int DX = Mo::DX, DQ = Mo::DQ, DY = Mo::DY;
  __shared__ float s_red[32 * (1 + BF_MAXD)];
  __shared__ float s_gqg[BF_MAXD * BF_MAXD];
  const int M = a.M, T = a.T, b = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
  const bool active = tid < M;
  const size_t bm = (size_t)b * M + (active ? tid : 0);

  __shared__ float s_R[BF_MAXD * BF_MAXD];
  // noise covariances of a step into shared memory: F_q Q F_q^T (F_q constant for every registered model) and R.
  // Constant ones are set once; `_get_params` (inference.py:21, :337-340) indexes covariances with a leading T axis.
  auto set_noise = [&](const float* Qsrc, const float* Rsrc) {       // NULL: keep
    if (tid == 0 && Qsrc) {
      float x0[DX], Fq[DX * DQ], Ql[DQ * DQ], FQ[DX * DQ];
#pragma unroll
      for (int i = 0; i < DX; ++i) x0[i] = mc.m0[i];
#pragma unroll
      for (int i = 0; i < DQ * DQ; ++i) Ql[i] = Qsrc[i];
      Mo::Fq(x0, mc.q0, 0.0f, mc, Fq);
      mm<DX, DQ, DQ>(Fq, Ql, FQ);
      float G[DX * DX];
      mmt<DX, DQ, DX>(FQ, Fq, G);
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) s_gqg[i] = G[i];
    }
    if (Rsrc)
      for (int i = tid; i < DY * DY; i += blockDim.x) s_R[i] = Rsrc[i];
  };
  set_noise(mc.Q, mc.R);
  __syncthreads();
  float GQG[DX * DX];
#pragma unroll
  for (int i = 0; i < DX * DX; ++i) GQG[i] = s_gqg[i];

  // carry: (weights, pred_means, pred_covs)   inference.py:367-369
  float w, m[DX], P[DX * DX];
  if (a.resume) {
    w = active ? a.carry_w[bm] : 0.0f;
#pragma unroll
    for (int i = 0; i < DX; ++i) m[i] = a.carry_m[bm * DX + i];
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) P[i] = a.carry_P[bm * DX * DX + i];
  } else {
    w = active ? (a.init_weights ? a.init_weights[bm] : 1.0f / (float)M) : 0.0f;
#pragma unroll
    for (int i = 0; i < DX; ++i) m[i] = a.init_means[bm * DX + i];
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) P[i] = a.init_covs ? a.init_covs[bm * DX * DX + i] : mc.P0[i];
  }

  const float jitter = (a.flags & BF_GSF_NO_JITTER) ? 0.0f : 1e-6f;
  const bool cov_plus_p = (a.flags & BF_GSF_COV_PLUS_P) != 0;
  // legacy order (gausssumfilt.py:51-70): the carry entering the call holds FILTERED moments -> predict first,
  // with the input of the first step
  if ((a.flags & BF_GSF_PREDICT_FIRST) && !a.resume) ekf_predict<Mo>(m, P, a.inputs ? a.inputs[0] : 0.0f, mc, GQG, cov_plus_p);

  for (int t = 0; t < T; ++t) {
    const float* yp = a.emissions + ((size_t)(a.emissions_per_trial ? b : 0) * T + t) * DY;
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = yp[k];
    const float u = a.inputs ? a.inputs[t] : 0.0f;
    const size_t btm = ((size_t)b * T + t) * M + (active ? tid : 0);
    if (TV && (a.Q_t || a.R_t)) {      // block-uniform; both the update and the predict of step t use index t
      __syncthreads();
      set_noise(a.Q_t ? a.Q_t + (size_t)t * DQ * DQ : nullptr, a.R_t ? a.R_t + (size_t)t * DY * DY : nullptr);
      __syncthreads();
      if (a.Q_t) {
#pragma unroll
        for (int i = 0; i < DX * DX; ++i) GQG[i] = s_gqg[i];
      }
    }

    float ll;
    if constexpr (TV) ll = ekf_condition<Mo>(m, P, y, u, mc, jitter, s_R);
    else ll = ekf_condition<Mo>(m, P, y, u, mc, jitter, mc.R);
    if (!active) ll = -INFINITY;
    if (active) {
      if (a.out_means) store_vec<DX>(a.out_means + btm * DX, m);
      if (a.out_covs) store_vec<DX * DX>(a.out_covs + btm * DX * DX, P);
    }
    // lls -= max ; w = exp(lls) * w ; w /= sum(w)       (inference.py:347-350)
    float mx = warp_max(ll);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = warp_max(lane < nwarps ? s_red[lane] : -INFINITY);
    float e = active ? expf(ll - mx) * w : 0.0f;
    float z = warp_sum(e);
    __syncthreads();
    if (lane == 0) s_red[warp] = z;
    __syncthreads();
    const float Z = warp_sum(lane < nwarps ? s_red[lane] : 0.0f);
    w = e / Z;
    if (active && a.out_weights) a.out_weights[btm] = w;
    if (tid == 0 && a.out_loglik) a.out_loglik[(size_t)b * T + t] = mx + logf(Z);
    if (a.out_mix_mean) {
      float v[DX];
#pragma unroll
      for (int k = 0; k < DX; ++k) v[k] = warp_sum(active ? w * m[k] : 0.0f);
      __syncthreads();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < DX; ++k) s_red[warp * DX + k] = v[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DX; ++k) v[k] = warp_sum(lane < nwarps ? s_red[lane * DX + k] : 0.0f);
      if (tid == 0) {
#pragma unroll
        for (int k = 0; k < DX; ++k) a.out_mix_mean[((size_t)b * T + t) * DX + k] = v[k];
      }
    }
    __syncthreads();                  // s_red reuse in the next step

    ekf_predict<Mo>(m, P, u, mc, GQG, cov_plus_p);
    if (active) {
      if (a.out_pred_means) store_vec<DX>(a.out_pred_means + btm * DX, m);
      if (a.out_pred_covs) store_vec<DX * DX>(a.out_pred_covs + btm * DX * DX, P);
    }
  }
  if (active) {
    a.carry_w[bm] = w;
#pragma unroll
    for (int i = 0; i < DX; ++i) a.carry_m[bm * DX + i] = m[i];
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) a.carry_P[bm * DX * DX + i] = P[i];
  }
}


// ---------------------------------------------------------------------------------------------------------
// Same filter, ONE WARP per trial, CPL components per lane (component c = lane + 32 j).  For the small state
// dimensions of the reference's experiments a component's (w, mean, cov) is a handful of registers, so a lane can
// carry several: the max and the sum across the mixture are then plain warp shuffles (no shared memory, no CTA
// barrier), amortised over CPL components, and the CPL independent EKFs give each thread instruction-level
// parallelism.  A CTA = 4 warps = 4 trials.  Used when M <= 32 CPL fits the register budget (bfilt_api.cu).
template <class Mo, int CPL>
__global__ void __launch_bounds__(128)
k_gsf_warp(const __grid_constant__ ModelConst mc, const __grid_constant__ GsfArgs a, int B) {
  constexpr int DX = Mo::DX, DQ = Mo::DQ, DY = Mo::DY;
  const int M = a.M, T = a.T, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;                                   // whole warp; no CTA-level barrier below

  float GQG[DX * DX];                                   // F_q Q F_q^T, F_q constant for every registered model
  {
    float x0[DX], Fq[DX * DQ], FQ[DX * DQ];
#pragma unroll
    for (int i = 0; i < DX; ++i) x0[i] = mc.m0[i];
    Mo::Fq(x0, mc.q0, 0.0f, mc, Fq);
    mm<DX, DQ, DQ>(Fq, mc.Q, FQ);
    mmt<DX, DQ, DX>(FQ, Fq, GQG);
  }
  const float jitter = (a.flags & BF_GSF_NO_JITTER) ? 0.0f : 1e-6f;
  const bool cov_plus_p = (a.flags & BF_GSF_COV_PLUS_P) != 0;

  float w[CPL], m[CPL][DX], P[CPL][DX * DX];
  bool act[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    act[j] = c < M;
    const size_t bm = (size_t)b * M + (act[j] ? c : 0);
    if (a.resume) {
      w[j] = act[j] ? a.carry_w[bm] : 0.0f;
#pragma unroll
      for (int i = 0; i < DX; ++i) m[j][i] = a.carry_m[bm * DX + i];
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) P[j][i] = a.carry_P[bm * DX * DX + i];
    } else {
      w[j] = act[j] ? (a.init_weights ? a.init_weights[bm] : 1.0f / (float)M) : 0.0f;
#pragma unroll
      for (int i = 0; i < DX; ++i) m[j][i] = a.init_means[bm * DX + i];
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) P[j][i] = a.init_covs ? a.init_covs[bm * DX * DX + i] : mc.P0[i];
      if (a.flags & BF_GSF_PREDICT_FIRST) ekf_predict<Mo>(m[j], P[j], a.inputs ? a.inputs[0] : 0.0f, mc, GQG, cov_plus_p);
    }
  }

  const float* yrow = a.emissions + (size_t)(a.emissions_per_trial ? b : 0) * T * DY;
  for (int t = 0; t < T; ++t) {
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = yrow[(size_t)t * DY + k];
    const float u = a.inputs ? a.inputs[t] : 0.0f;
    const size_t bt = (size_t)b * T + t;

    float ll[CPL], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      ll[j] = ekf_condition<Mo>(m[j], P[j], y, u, mc, jitter, mc.R);
      if (!act[j]) ll[j] = -INFINITY;
      mx = fmaxf(mx, ll[j]);
      if (act[j]) {
        const size_t btm = bt * M + lane + 32 * j;
        if (a.out_means) store_vec<DX>(a.out_means + btm * DX, m[j]);
        if (a.out_covs) store_vec<DX * DX>(a.out_covs + btm * DX * DX, P[j]);
      }
    }
    // lls -= max ; w = exp(lls) * w ; w /= sum(w)       (inference.py:347-350)
    mx = warp_max(mx);
    float z = 0.0f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) { w[j] = act[j] ? expf(ll[j] - mx) * w[j] : 0.0f; z += w[j]; }
    const float Z = warp_sum(z);
    const float rZ = 1.0f / Z;
    float mix[DX];
#pragma unroll
    for (int k = 0; k < DX; ++k) mix[k] = 0.0f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      w[j] = __fmul_rn(w[j], rZ);
      if (act[j] && a.out_weights) a.out_weights[bt * M + lane + 32 * j] = w[j];
#pragma unroll
      for (int k = 0; k < DX; ++k) mix[k] += w[j] * m[j][k];
    }
    if (lane == 0 && a.out_loglik) a.out_loglik[bt] = mx + logf(Z);
    if (a.out_mix_mean) {
#pragma unroll
      for (int k = 0; k < DX; ++k) mix[k] = warp_sum(mix[k]);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < DX; ++k) a.out_mix_mean[bt * DX + k] = mix[k];
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      ekf_predict<Mo>(m[j], P[j], u, mc, GQG, cov_plus_p);
      if (act[j]) {
        const size_t btm = bt * M + lane + 32 * j;
        if (a.out_pred_means) store_vec<DX>(a.out_pred_means + btm * DX, m[j]);
        if (a.out_pred_covs) store_vec<DX * DX>(a.out_pred_covs + btm * DX * DX, P[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    if (act[j]) {
      const size_t bm = (size_t)b * M + lane + 32 * j;
      a.carry_w[bm] = w[j];
#pragma unroll
      for (int i = 0; i < DX; ++i) a.carry_m[bm * DX + i] = m[j][i];
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) a.carry_P[bm * DX * DX + i] = P[j][i];
    }
  }
}
