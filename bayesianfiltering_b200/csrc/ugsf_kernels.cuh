// ugsf_kernels.cuh -- unscented Gaussian-sum filter, persistent over time (sm_100a).
//
// Restates gaussfiltax/inference.py:379-456 (`unscented_gaussian_sum_filter`): per component
// `_ukf_condition_on_nonadditive` (:198-224) on the carried prediction, the max-shifted evidence reweighting across the M
// components (:424-427), then `_ukf_predict_nonadditive` (:146-174); sigma points from `utils._get_sigma_points`
// (utils.py:247-254).  Same skeleton as k_gsf (gsf_kernels.cuh): one CTA per independent trial, one thread per mixture
// component, a component's (w, mean, cov) in that thread's registers for all T steps.
//
// The unscented transform of the AUGMENTED Gaussian N([m; n0], blockdiag(P, N)) with n = DX + DN:
//   lambda = alpha^2 (n + kappa) - n,  c = sqrt(n + lambda),  L = sqrtm(blockdiag(P, N)) = blockdiag(sqrtm(P), sqrtm(N))
//   2n sigma points  [m; n0] +- c L[:, i]  -- no centre point; the centre enters through g(m, n0) below
//   mu  = sum_i g_i / (2 (lambda + n)) + g(m, n0) lambda / (lambda + n)
//   S   = sum_i (g_i - mu)(g_i - mu)^T / (2 (lambda + n)) + (lambda / (lambda + n) + 1 - alpha^2 + beta) (g(m,n0) - mu)(.)^T
//   C   = sum_i (g_i - mu)(x_i - m)^T / (2 (lambda + n))                       (update only; x_i = state part)
// `scipy.linalg.sqrtm` of a symmetric PSD matrix is its principal square root V diag(sqrt(lambda)) V^T; here: cyclic Jacobi in
// fp32 on the symmetrised matrix, negative eigenvalues (an fp32 covariance that lost definiteness) clamped to 0 -- the
// real part of the reference's complex root.  sqrtm(Q), sqrtm(R) are constants of the run, computed on the host in float64.
// Update: K = solve(S + 1e-6, C)^T (psd_solve, LU), P+ = P - K S K^T, m+ = m + K (y - mu), ll = MVN(mu, S).log_prob(y).
#pragma once
#include "gsf_kernels.cuh"
#include "agsf_kernels.cuh"      // chol_lower

// mode 0: the reference's unscented_gaussian_sum_filter (noise-augmented sigma points, sqrtm).
// mode 1: legacy gaussfilt.UKF.moment_approx (gaussfilt.py:146-170): ADDITIVE noise, 2 dx + 1 sigma points from the Cholesky
//         factor of P (utils.split_to_sigma_points, utils.py:52-65), noise covariance added to the transformed covariance.
// mode 2: legacy gaussfilt.MCF.moment_approx (gaussfilt.py:186-198): moments from `np` Monte-Carlo draws of N(m, P).
struct UkfArgs {
  float alpha, beta, kappa;
  const float* sqrtQ;       // [DQ,DQ] principal square root of Q (device)          (mode 0)
  const float* sqrtR;       // [DR,DR]
  int mode;
  int np;                   // mode 2: draws per moment approximation
  const float* draws;       // mode 2: [B,T,M,2,np,DX] coloured offsets x - m (prediction draws, update draws), or NULL (Philox)
  unsigned long long seed;
};

// principal square root of a symmetric PSD N x N matrix: cyclic Jacobi, every index compile-time (registers).  The sweeps run
// in float64: the covariances of the tracking models span five orders of magnitude (position 1e-1, velocity 1e-6) and a
// float32 eigen-decomposition loses the small eigen-directions to ~1e-3 relative (measured against the reference fixtures);
// N <= 4 here, so the float64 pipe is not a cost that matters.
template <int N>
__device__ __forceinline__ void sym_sqrt(const float* P, float* L) {
  if constexpr (N == 1) { L[0] = sqrtf(fmaxf(P[0], 0.0f)); return; }
  double A[N * N], V[N * N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) { A[i * N + j] = 0.5 * ((double)P[i * N + j] + (double)P[j * N + i]); V[i * N + j] = (i == j) ? 1.0 : 0.0; }
  constexpr int SWEEPS = (N == 2) ? 2 : (N <= 4 ? 8 : 12);
#pragma unroll 1
  for (int sw = 0; sw < SWEEPS; ++sw) {
#pragma unroll
    for (int p = 0; p < N - 1; ++p)
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        const double app = A[p * N + p], aqq = A[q * N + q];
        // rotation that annihilates A[p][q]; skipped when it is already negligible against the diagonal
        const bool rot = fabs(apq) > 1e-300 && fabs(apq) > 1e-22 * (fabs(app) + fabs(aqq));
        const double th = (aqq - app) / (2.0 * (rot ? apq : 1.0));
        const double tt = copysign(1.0, th) / (fabs(th) + sqrt(fma(th, th, 1.0)));
        const double cs = rot ? rsqrt(fma(tt, tt, 1.0)) : 1.0;
        const double sn = rot ? tt * cs : 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {          // A <- A J   (columns p, q)
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = cs * akp - sn * akq; A[k * N + q] = sn * akp + cs * akq;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {          // A <- J^T A (rows p, q)
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = cs * apk - sn * aqk; A[q * N + k] = sn * apk + cs * aqk;
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {          // V <- V J
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = cs * vkp - sn * vkq; V[k * N + q] = sn * vkp + cs * vkq;
        }
      }
  }
  double rt[N];
#pragma unroll
  for (int i = 0; i < N; ++i) rt[i] = sqrt(fmax(A[i * N + i], 0.0));
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(V[i * N + k] * rt[k], V[j * N + k], s);
      L[i * N + j] = (float)s;
    }
}

// Unscented transform of g over N([m; n0], blockdiag(P, Nn)); Lx = sqrtm(P) [DX,DX], Ln = sqrtm(Nn) [DN,DN] (global).
// g(x[DX], nz[DN], out[DO]).  Returns mu [DO], S [DO,DO] and, if WITH_C, C [DO,DX].
template <int DX, int DN, int DO, bool WITH_C, class G>
__device__ __forceinline__ void unscented(const float* m, const float* Lx, const float* n0, const float* Ln, float alpha, float beta,
                                          float kappa, G g, float* mu, float* S, float* C) {
  constexpr int n = DX + DN;
  const float lam = alpha * alpha * ((float)n + kappa) - (float)n;          // :162 / :208
  const float c = sqrtf((float)n + lam);
  const float ws = 1.0f / (2.0f * (lam + (float)n));
  const float wc = lam / (lam + (float)n);
  const float wcc = wc + 1.0f - alpha * alpha + beta;
  constexpr int DNA = DN > 0 ? DN : 1;
  float g0[DO], pts[2 * n][DO], nz0[DNA];
#pragma unroll
  for (int k = 0; k < DN; ++k) nz0[k] = n0[k];
  g(m, nz0, g0);
#pragma unroll
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
      float x[DX], nz[DNA];
      const float cc = sgn ? -c : c;
#pragma unroll
      for (int k = 0; k < DX; ++k) x[k] = (i < DX) ? fmaf(cc, Lx[k * DX + (i < DX ? i : 0)], m[k]) : m[k];
#pragma unroll
      for (int k = 0; k < DN; ++k) nz[k] = (i >= DX) ? fmaf(cc, Ln[k * DN + (i >= DX ? i - DX : 0)], nz0[k]) : nz0[k];
      g(x, nz, pts[2 * i + sgn]);
    }
  }
#pragma unroll
  for (int a = 0; a < DO; ++a) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 2 * n; ++i) s += pts[i][a];
    mu[a] = fmaf(s, ws, g0[a] * wc);
  }
#pragma unroll
  for (int a = 0; a < DO; ++a)
#pragma unroll
    for (int b = 0; b < DO; ++b) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 2 * n; ++i) s = fmaf(pts[i][a] - mu[a], pts[i][b] - mu[b], s);
      S[a * DO + b] = fmaf(s, ws, wcc * (g0[a] - mu[a]) * (g0[b] - mu[b]));
    }
  if constexpr (WITH_C) {
#pragma unroll
    for (int a = 0; a < DO; ++a)
#pragma unroll
      for (int k = 0; k < DX; ++k) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < DX; ++i) {           // x_i - m of the ROUNDED sigma points (:221); 0 for the noise points
          const float dp = fmaf(c, Lx[k * DX + i], m[k]) - m[k], dm = fmaf(-c, Lx[k * DX + i], m[k]) - m[k];
          s = fmaf(pts[2 * i][a] - mu[a], dp, s);
          s = fmaf(pts[2 * i + 1][a] - mu[a], dm, s);
        }
        C[a * DX + k] = s * ws;
      }
  }
}

// Gain, posterior and evidence from the moments (mu, S, C) of a transform: X holds C [DY,DX] on entry.
//   K = solve(S + jitter, C)^T, P+ = P - K S K^T, m+ = m + K (y - mu), ll = MVN(mu, S).log_prob(y)
template <int DX, int DY>
__device__ __forceinline__ float kalman_from_moments(float* m, float* P, const float* y, const float* mu, const float* S, float* X, float jitter) {
  float Sj[DY * DY];
#pragma unroll
  for (int i = 0; i < DY * DY; ++i) Sj[i] = S[i] + jitter;      // psd_solve: + 1e-6 on EVERY element
  lu_solve<DY, DX>(Sj, X);                                     // X = solve(S', C);  K = X^T   (:222)
  float KS[DX * DY];
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
  for (int i = 0; i < DX; ++i)
#pragma unroll
    for (int j = 0; j < DX; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < DY; ++k) s += KS[i * DY + k] * X[k * DX + j];
      P[i * DX + j] -= s;                                      // P - K S K^T   (:223)
    }
  float d[DY];
#pragma unroll
  for (int i = 0; i < DY; ++i) d[i] = y[i] - mu[i];
#pragma unroll
  for (int i = 0; i < DX; ++i) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < DY; ++k) s += X[k * DX + i] * d[k];
    m[i] += s;                                                 // m + K (y - mu)   (:224)
  }
  return mvn_logpdf_chol<DY>(S, d);
}

// Monte-Carlo moments of g over N(m, P) (gaussfilt.MCF.moment_approx, gaussfilt.py:186-198): particles m + d_i with the
// offsets d_i supplied (`dr` [np,DX]) or drawn here (Cholesky of P times Philox normals); two passes, g evaluated twice
// rather than np transformed particles kept per thread.  Returns mu, S = cov + sample covariance, C [DO,DX] (if WITH_C).
template <int DX, int DO, bool WITH_C, class G>
__device__ __forceinline__ void mc_moments(const float* m, const float* P, const float* cov, int np, const float* dr,
                                           unsigned long long seed, uint32_t stream, uint32_t t, uint32_t b, uint32_t comp, G g,
                                           float* mu, float* S, float* C) {
  float L[DX * DX];
  if (!dr) chol_lower<DX>(P, L);
  auto particle = [&](int i, float* x) {
    if (dr) {
#pragma unroll
      for (int k = 0; k < DX; ++k) x[k] = m[k] + dr[(size_t)i * DX + k];
    } else {
      float z[DX];
      philox_normals<DX>(seed, stream, t, b, ((uint64_t)comp << 32) | (uint32_t)i, z);
#pragma unroll
      for (int k = 0; k < DX; ++k) {
        float s2 = m[k];
#pragma unroll
        for (int q = 0; q <= k; ++q) s2 = fmaf(L[k * DX + q], z[q], s2);
        x[k] = s2;
      }
    }
  };
  const float inp = 1.0f / (float)np;
#pragma unroll
  for (int a = 0; a < DO; ++a) mu[a] = 0.0f;
  for (int i = 0; i < np; ++i) {
    float x[DX], o[DO];
    particle(i, x);
    g(x, o);
#pragma unroll
    for (int a = 0; a < DO; ++a) mu[a] += o[a];
  }
#pragma unroll
  for (int a = 0; a < DO; ++a) mu[a] *= inp;
#pragma unroll
  for (int i = 0; i < DO * DO; ++i) S[i] = 0.0f;
  if constexpr (WITH_C) {
#pragma unroll
    for (int i = 0; i < DO * DX; ++i) C[i] = 0.0f;
  }
  for (int i = 0; i < np; ++i) {
    float x[DX], o[DO];
    particle(i, x);
    g(x, o);
#pragma unroll
    for (int a = 0; a < DO; ++a)
#pragma unroll
      for (int c = 0; c < DO; ++c) S[a * DO + c] = fmaf(o[a] - mu[a], o[c] - mu[c], S[a * DO + c]);
    if constexpr (WITH_C) {
#pragma unroll
      for (int a = 0; a < DO; ++a)
#pragma unroll
        for (int k = 0; k < DX; ++k) C[a * DX + k] = fmaf(o[a] - mu[a], x[k] - m[k], C[a * DX + k]);
    }
  }
#pragma unroll
  for (int i = 0; i < DO * DO; ++i) S[i] = fmaf(S[i], inp, cov[i]);
  if constexpr (WITH_C) {
#pragma unroll
    for (int i = 0; i < DO * DX; ++i) C[i] *= inp;
  }
}

// where the Monte-Carlo draws of (trial b, step t, component c, phase ph) start, or NULL
__device__ __forceinline__ const float* mc_draws(const UkfArgs& uk, int b, int T, int t, int M, int c, int ph, int dx) {
  if (!uk.draws || t >= T) return nullptr;
  return uk.draws + (((((size_t)b * T + t) * M + c) * 2 + ph) * uk.np) * dx;
}

// update of one component: `_ukf_condition_on_nonadditive` (inference.py:198-224) or the legacy moment approximations
template <class Mo>
__device__ __forceinline__ float ukf_condition(float* m, float* P, const float* y, float u, const ModelConst& mc, const UkfArgs& uk,
                                               float jitter, const float* dr, int t, int b, int comp) {
  constexpr int DX = Mo::DX, DY = Mo::DY, DR = Mo::DR;
  float L[DX * DX], mu[DY], S[DY * DY], X[DY * DX];
  if (uk.mode == 0) {
    sym_sqrt<DX>(P, L);
    unscented<DX, DR, DY, true>(m, L, mc.r0, uk.sqrtR, uk.alpha, uk.beta, uk.kappa,
                                [&](const float* x, const float* r, float* o) { Mo::h(x, r, u, mc, o); }, mu, S, X);
  } else if (uk.mode == 1) {         // gaussfilt.UKF, 'upd': g(x) = h(x, 0, 0), + R
    chol_lower<DX>(P, L);
    float none[1] = {0.0f};
    unscented<DX, 0, DY, true>(m, L, none, none, uk.alpha, uk.beta, uk.kappa,
                               [&](const float* x, const float*, float* o) { float r0[DR]; for (int k = 0; k < DR; ++k) r0[k] = 0.0f; Mo::h(x, r0, u, mc, o); },
                               mu, S, X);
#pragma unroll
    for (int i = 0; i < DY * DY; ++i) S[i] += mc.R[i];
  } else {                           // gaussfilt.MCF, 'upd'
    mc_moments<DX, DY, true>(m, P, mc.R, uk.np, dr, uk.seed, 32u, (uint32_t)t, (uint32_t)b, (uint32_t)comp,
                             [&](const float* x, float* o) { float r0[DR]; for (int k = 0; k < DR; ++k) r0[k] = 0.0f; Mo::h(x, r0, u, mc, o); },
                             mu, S, X);
  }
  return kalman_from_moments<DX, DY>(m, P, y, mu, S, X, jitter);
}

// prediction of one component: `_ukf_predict_nonadditive` (inference.py:146-174) or the legacy moment approximations ('pred')
template <class Mo>
__device__ __forceinline__ void ukf_predict(float* m, float* P, float u, const ModelConst& mc, const UkfArgs& uk, const float* dr,
                                            int t, int b, int comp) {
  constexpr int DX = Mo::DX, DQ = Mo::DQ;
  float L[DX * DX], mu[DX], S[DX * DX];
  if (uk.mode == 0) {
    sym_sqrt<DX>(P, L);
    unscented<DX, DQ, DX, false>(m, L, mc.q0, uk.sqrtQ, uk.alpha, uk.beta, uk.kappa,
                                 [&](const float* x, const float* q, float* o) { Mo::f(x, q, u, mc, o); }, mu, S, nullptr);
  } else {
    // additive-noise legacy classes: f(x) = f(x, 0, 0), + Q (needs F_q = I, i.e. DQ == DX: checked on the host)
    auto fx = [&](const float* x, float* o) { float q0[DQ]; for (int k = 0; k < DQ; ++k) q0[k] = 0.0f; Mo::f(x, q0, u, mc, o); };
    if (uk.mode == 1) {
      chol_lower<DX>(P, L);
      float none[1] = {0.0f};
      unscented<DX, 0, DX, false>(m, L, none, none, uk.alpha, uk.beta, uk.kappa, [&](const float* x, const float*, float* o) { fx(x, o); }, mu, S, nullptr);
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) S[i] += mc.Q[(DQ == DX) ? i : 0];
    } else {
      mc_moments<DX, DX, false>(m, P, mc.Q, uk.np, dr, uk.seed, 31u, (uint32_t)t, (uint32_t)b, (uint32_t)comp, fx, mu, S, nullptr);
    }
  }
#pragma unroll
  for (int i = 0; i < DX; ++i) m[i] = mu[i];
#pragma unroll
  for (int i = 0; i < DX * DX; ++i) P[i] = S[i];
}

template <class Mo>
__global__ void k_ugsf(const __grid_constant__ ModelConst mc, const __grid_constant__ GsfArgs a, const __grid_constant__ UkfArgs uk) {
  constexpr int DX = Mo::DX, DY = Mo::DY;
  __shared__ float s_red[32 * (1 + BF_MAXD)];
  const int M = a.M, T = a.T, b = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
  const bool active = tid < M;
  const size_t bm = (size_t)b * M + (active ? tid : 0);
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
  const int comp = active ? tid : 0;
  // legacy order (gaussfilt.py:108-113): the carry entering the call holds FILTERED moments -> predict first
  if ((a.flags & BF_GSF_PREDICT_FIRST) && !a.resume)
    ukf_predict<Mo>(m, P, a.inputs ? a.inputs[0] : 0.0f, mc, uk, mc_draws(uk, b, T, 0, M, comp, 0, DX), 0, b, comp);
  for (int t = 0; t < T; ++t) {
    const float* yp = a.emissions + ((size_t)(a.emissions_per_trial ? b : 0) * T + t) * DY;
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = yp[k];
    const float u = a.inputs ? a.inputs[t] : 0.0f;
    const size_t btm = ((size_t)b * T + t) * M + (active ? tid : 0);
    float ll = ukf_condition<Mo>(m, P, y, u, mc, uk, jitter, mc_draws(uk, b, T, t, M, comp, 1, DX), t, b, comp);
    if (!active) ll = -INFINITY;
    if (active) {
      if (a.out_means) store_vec<DX>(a.out_means + btm * DX, m);
      if (a.out_covs) store_vec<DX * DX>(a.out_covs + btm * DX * DX, P);
    }
    // lls -= max ; w = exp(lls) * w ; w /= sum(w)       (inference.py:424-427)
    float mx = warp_max(ll);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = warp_max(lane < nwarps ? s_red[lane] : -INFINITY);
    const float e = active ? expf(ll - mx) * w : 0.0f;
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
    __syncthreads();
    {   // predict-first order: this is the prediction of step t + 1 (its draws, if supplied, are those of t + 1)
      const bool pf = (a.flags & BF_GSF_PREDICT_FIRST) != 0;
      ukf_predict<Mo>(m, P, u, mc, uk, mc_draws(uk, b, T, pf ? t + 1 : t, M, comp, 0, DX), pf ? t + 1 : t, b, comp);
    }
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
