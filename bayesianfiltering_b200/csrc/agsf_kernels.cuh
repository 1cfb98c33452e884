// agsf_kernels.cuh -- augmented Gaussian-sum filter ("speedy" form), persistent over time (sm_100a).
//
// Restates gaussfiltax/inference.py:621-812 (`speedy_augmented_gaussian_sum_filter`).  One step of a mixture of M
// Gaussians N(m_c, P_c), weights 1/M:
//   branch 1   Delta_c = a0 P_c ;  z_cj = m_c + chol(P_c - Delta_c) xi_cj,  j < N1          (:666, :686-692)
//   predict    (mu_cj, Pp_cj) = _predict(z_cj, Delta_c)   -- EKF prediction of N(z_cj, Delta_c)       (:702-703, :51-70)
//   branch 2   Lambda_cj = a1 Pp_cj ;  s_cjk = mu_cj + chol(Pp_cj - Lambda_cj) zeta_cjk,  k < N2     (:718, :728-734)
//   update     (ll, m', P')_cjk = _condition_on(s_cjk, Lambda_cj)                          (:741-742, :72-105)
//   weights    w_cjk = exp(ll - max) * ((1/M) / N1) / N2, normalised over the L = M N1 N2 leaves      (:743-749)
//   resample   idx = jr.choice(key, L, (M,), p=w): cumsum in associative-scan order, r = total (1 - u),
//              fixed-bisection searchsorted; the M survivors get weight 1/M                 (:764-768)
// One CTA per independent trial; a thread owns a leaf (c, j, k) and recomputes the branch-1 / predict part of its parent
// (N2-fold redundant, a few hundred flops) rather than synchronise on it.  The L updated components go to a scratch in
// global memory (L2-resident) from which the M survivors are gathered; weights, cdf and the carry live in shared memory.
// The resampler is the particle filter's: canonical tree cumsum (scan.cuh) + the reference's bisection, bit-exact ancestors.
// Randomness: xi [M,DX,N1], zeta [M*N1,DX,N2] standard normals and u [M] uniforms, supplied (one set reused at every step
// -- what the reference does, its keys being closure constants -- or one per step) or drawn from Philox.
#pragma once
#include "gsf_kernels.cuh"
#include "scan.cuh"

#define AGSF_WARPS 8

struct AgsfArgs {
  int M, N1, N2, T, emissions_per_trial, rand_per_step;
  float a0, a1;
  unsigned long long seed;
  const float* emissions; const float* inputs;
  const float* init_means;                 // [B,M,DX]
  const float* init_covs;                  // [B,M,DX,DX] or NULL (P0)
  const float* z_iid; const float* s_iid; const float* res_u;     // [(T),M,DX,N1] [(T),M*N1,DX,N2] [(T),M] or NULL (Philox)
  float* leaf_m; float* leaf_P;            // scratch [B,L,DX], [B,L,DX,DX]
  float* out_weights; float* out_means; float* out_covs;          // [B,T,M] [B,T,M,DX] [B,T,M,DX,DX]
  float* out_loglik;                       // [B,T]   max + log sum (evidence of the step under the branched mixture)
  float* out_pre_weights;                  // [B,T,L] normalised leaf weights before resampling (aux "pre_weights")
  float* out_updated_means;                // [B,T,L,DX]  (aux "updated_means")
  int* out_ancestors;                      // [B,T,M] leaf index of each survivor
};

// lower Cholesky factor of a symmetric matrix, registers; a non-positive pivot gives NaN like jnp.linalg.cholesky
template <int N>
__device__ __forceinline__ void chol_lower(const float* A, float* L) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j > i) { L[i * N + j] = 0.0f; continue; }
      float s = A[i * N + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fmaf(-L[i * N + k], L[j * N + k], s);
      L[i * N + j] = (i == j) ? sqrtf(s) : s / L[j * N + j];
    }
}

template <class Mo>
__global__ void __launch_bounds__(AGSF_WARPS * 32)
k_agsf(const __grid_constant__ ModelConst mc, const __grid_constant__ AgsfArgs a) {
  constexpr int DX = Mo::DX, DQ = Mo::DQ, DY = Mo::DY, NT = AGSF_WARPS * 32;
  extern __shared__ __align__(16) float s_dyn[];
  const int M = a.M, N1 = a.N1, N2 = a.N2, L = M * N1 * N2, T = a.T, b = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  float* s_m = s_dyn;                          // carry means [M,DX]
  float* s_P = s_m + M * DX;                   // carry covariances [M,DX,DX]
  float* s_w = s_P + M * DX * DX;              // leaf log-likelihoods, then normalised weights [L]
  float* s_E = s_w + ((L + 4) & ~3);           // exclusive canonical cdf [L+1]
  float* s_scan = s_E + ((L + 5) & ~3);        // scan scratch
  __shared__ float s_red[AGSF_WARPS];
  __shared__ float s_gqg[BF_MAXD * BF_MAXD];
  if (tid == 0) {                              // F_q Q F_q^T, F_q constant for every registered model
    float x0[DX], Fq[DX * DQ], FQ[DX * DQ], G[DX * DX];
#pragma unroll
    for (int i = 0; i < DX; ++i) x0[i] = mc.m0[i];
    Mo::Fq(x0, mc.q0, 0.0f, mc, Fq);
    mm<DX, DQ, DQ>(Fq, mc.Q, FQ);
    mmt<DX, DQ, DX>(FQ, Fq, G);
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) s_gqg[i] = G[i];
  }
  for (int i = tid; i < M * DX; i += NT) s_m[i] = a.init_means[(size_t)b * M * DX + i];
  for (int i = tid; i < M * DX * DX; i += NT) s_P[i] = a.init_covs ? a.init_covs[(size_t)b * M * DX * DX + i] : mc.P0[i % (DX * DX)];
  __syncthreads();
  float GQG[DX * DX];
#pragma unroll
  for (int i = 0; i < DX * DX; ++i) GQG[i] = s_gqg[i];
  float* lm = a.leaf_m + (size_t)b * L * DX;
  float* lP = a.leaf_P + (size_t)b * L * DX * DX;
  const float invM = 1.0f / (float)M;
  const int levels = bisect_levels(L);

  for (int t = 0; t < T; ++t) {
    const float* yp = a.emissions + ((size_t)(a.emissions_per_trial ? b : 0) * T + t) * DY;
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = yp[k];
    const float u = a.inputs ? a.inputs[t] : 0.0f;
    const int tr = a.rand_per_step ? t : 0;
    float lmax = -INFINITY;
    for (int leaf = tid; leaf < L; leaf += NT) {
      const int c = leaf / (N1 * N2), j = (leaf / N2) % N1, k = leaf % N2, cj = c * N1 + j;
      float m[DX], P[DX * DX], D[DX * DX], A[DX * DX], Lc[DX * DX], z[DX];
#pragma unroll
      for (int i = 0; i < DX; ++i) m[i] = s_m[c * DX + i];
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) { P[i] = s_P[c * DX * DX + i]; D[i] = a.a0 * P[i]; A[i] = P[i] - D[i]; }
      chol_lower<DX>(A, Lc);                                                        // :686
#pragma unroll
      for (int i = 0; i < DX; ++i) {                                                // xi[c, i, j]
        z[i] = a.z_iid ? a.z_iid[(((size_t)tr * M + c) * DX + i) * N1 + j] : 0.0f;
      }
      if (!a.z_iid) {
        float zz[DX];
        philox_normals<DX>(a.seed, 11u, (uint32_t)(a.rand_per_step ? t : 0), (uint32_t)b, (uint64_t)cj, zz);
#pragma unroll
        for (int i = 0; i < DX; ++i) z[i] = zz[i];
      }
#pragma unroll
      for (int i = 0; i < DX; ++i) {                                                // m + L1 xi   (:688-691)
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q <= i; ++q) s = fmaf(Lc[i * DX + q], z[q], s);
        m[i] += s;
      }
      ekf_predict<Mo>(m, D, u, mc, GQG);                                            // (m, D) <- _predict(z, Delta)   (:702-703)
#pragma unroll
      for (int i = 0; i < DX * DX; ++i) { P[i] = a.a1 * D[i]; A[i] = D[i] - P[i]; } // Lambda = a1 Pp   (:718)
      chol_lower<DX>(A, Lc);                                                        // :728
#pragma unroll
      for (int i = 0; i < DX; ++i) z[i] = a.s_iid ? a.s_iid[(((size_t)tr * M * N1 + cj) * DX + i) * N2 + k] : 0.0f;
      if (!a.s_iid) {
        float zz[DX];
        philox_normals<DX>(a.seed, 12u, (uint32_t)(a.rand_per_step ? t : 0), (uint32_t)b, (uint64_t)leaf, zz);
#pragma unroll
        for (int i = 0; i < DX; ++i) z[i] = zz[i];
      }
#pragma unroll
      for (int i = 0; i < DX; ++i) {                                                // mu + L2 zeta   (:730-733)
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q <= i; ++q) s = fmaf(Lc[i * DX + q], z[q], s);
        m[i] += s;
      }
      const float ll = ekf_condition<Mo>(m, P, y, u, mc, 1e-6f, mc.R);               // :741-742
      s_w[leaf] = ll;
      lmax = fmaxf(lmax, ll);
      if (ll != ll) lmax = NAN;                                                     // jnp.max propagates NaN
      store_vec<DX>(lm + (size_t)leaf * DX, m);
      store_vec<DX * DX>(lP + (size_t)leaf * DX * DX, P);
      if (a.out_updated_means) store_vec<DX>(a.out_updated_means + (((size_t)b * T + t) * L + leaf) * DX, m);
    }
    // lls -= max ; weights = exp(lls) * ((w / N1) / N2) ; weights /= sum     (:745-749)
    {
      float x = lmax;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const float v = __shfl_xor_sync(BF_FULL, x, o); x = (x != x || v != v) ? NAN : fmaxf(x, v); }
      if (lane == 0) s_red[warp] = x;
    }
    __syncthreads();
    float mx = s_red[0];
#pragma unroll
    for (int i = 1; i < AGSF_WARPS; ++i) { const float v = s_red[i]; mx = (mx != mx || v != v) ? NAN : fmaxf(mx, v); }
    const float uw = (invM / (float)N1) / (float)N2;
    float zsum = 0.0f;
    for (int leaf = tid; leaf < L; leaf += NT) { const float e = expf(s_w[leaf] - mx) * uw; s_w[leaf] = e; zsum += e; }
    zsum = warp_sum(zsum);
    __syncthreads();
    if (lane == 0) s_red[warp] = zsum;
    __syncthreads();
    float Z = 0.0f;
#pragma unroll
    for (int i = 0; i < AGSF_WARPS; ++i) Z += s_red[i];
    for (int leaf = tid; leaf < L; leaf += NT) {
      const float w = s_w[leaf] / Z;
      s_w[leaf] = w;
      if (a.out_pre_weights) a.out_pre_weights[((size_t)b * T + t) * L + leaf] = w;
    }
    if (tid == 0 && a.out_loglik) a.out_loglik[(size_t)b * T + t] = mx + logf(Z);
    __syncthreads();
    // resampling: jr.choice(p = weights)  (:764)  -- canonical cumsum, r = total (1 - u), fixed bisection
    cta_exclusive_scan<AGSF_WARPS>(s_w, s_E, L, s_scan);
    const float total = s_E[L];
    for (int i = tid; i < M; i += NT) {
      const float uu = a.res_u ? a.res_u[(size_t)tr * M + i]
                               : philox_uniform(a.seed, 13u, (uint32_t)(a.rand_per_step ? t : 0), (uint32_t)b, (uint64_t)i);
      const float r = __fmul_rn(total, __fsub_rn(1.0f, uu));
      const int idx = bisect_left_exclusive(s_E, L, levels, r);
      const size_t o = ((size_t)b * T + t) * M + i;
      if (a.out_ancestors) a.out_ancestors[o] = idx;
      if (a.out_weights) a.out_weights[o] = invM;                                   // :768
#pragma unroll
      for (int q = 0; q < DX; ++q) {
        const float v = __ldcg(lm + (size_t)idx * DX + q);
        s_m[i * DX + q] = v;
        if (a.out_means) a.out_means[o * DX + q] = v;
      }
#pragma unroll
      for (int q = 0; q < DX * DX; ++q) {
        const float v = __ldcg(lP + (size_t)idx * DX * DX + q);
        s_P[i * DX * DX + q] = v;
        if (a.out_covs) a.out_covs[o * DX * DX + q] = v;
      }
    }
    __syncthreads();
  }
}

#ifdef BF_TU_MAIN
// ---------------------------------------------------------------------------------------------------------------------
// utils.optimal_resampling (gaussfiltax/utils.py:216-244; Fearnhead & Clifford 2003): from M weights keep the largest ones
// exactly and resample the small ones with a common threshold weight p, returning N (index, weight) pairs.  One CTA;
// M <= 4096.  u [M]: the uniforms of its jr.choice(key, M, (M,), p=res_weights).
#define OPTRES_MAX 4096
__global__ void __launch_bounds__(AGSF_WARPS * 32)
k_optimal_resampling(const float* __restrict__ w, int M, int N, const float* __restrict__ u, int* __restrict__ idx_out,
                     float* __restrict__ w_out) {
  constexpr int NT = AGSF_WARPS * 32;
  extern __shared__ __align__(16) float s_dyn[];
  int P2 = 1;
  while (P2 < M) P2 <<= 1;
  float* key = s_dyn;                               // [P2] sorted weights
  int* id = reinterpret_cast<int*>(key + P2);       // [P2] jnp.argsort (stable: ties keep index order)
  float* rw = reinterpret_cast<float*>(id + P2);    // [M+4]  resampling weights
  float* E = rw + ((M + 4) & ~3);                   // [M+1]  exclusive canonical cdf
  int* ridx = reinterpret_cast<int*>(E + ((M + 5) & ~3));   // [M]
  float* scratch = reinterpret_cast<float*>(ridx + ((M + 3) & ~3));
  __shared__ float s_p, s_tot;
  const int tid = threadIdx.x;
  for (int i = tid; i < P2; i += NT) { key[i] = (i < M) ? w[i] : INFINITY; id[i] = i; }
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)                 // bitonic sort, ascending by (weight, index)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += NT) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const float a = key[i], b = key[l];
          const int ia = id[i], ib = id[l];
          const bool gt = (a > b) || (a == b && ia > ib);
          if (gt == up) { key[i] = b; key[l] = a; id[i] = ib; id[l] = ia; }
        }
      }
      __syncthreads();
    }
  if (tid == 0) {
    // ps[k] = (sum of the M-N+k+1 smallest weights) / (k+1), k < N-1, flipped (:228-229); the one `ind` with
    // sorted[M-ind-1] < ps[ind-1] < sorted[M-ind] (:230-232) is the number of weights kept exactly
    int Lk = 0;
    double c = 0.0;
    for (int jn = 0; jn < M - N; ++jn) c += (double)key[jn];
    for (int k = 0; k < N - 1; ++k) {
      c += (double)key[M - N + k];
      const float psk = (float)c / (float)(k + 1);
      const int ind = N - 1 - k;                    // after the flip, ps'[ind-1] = ps[N-1-ind] = ps[k]
      if (key[M - ind - 1] < psk && psk < key[M - ind]) Lk += ind;
    }
    float p = 1.0f / (float)N;                      // :233
    if (Lk > 0) {
      double cc = 0.0;
      for (int jn = 0; jn <= M - 1 - Lk; ++jn) cc += (double)key[jn];      // ps'[L-1] = ps[N-1-L]: rows up to M-N+(N-1-L)
      p = (float)cc / (float)(N - Lk);
    }
    s_p = p;
    float tot = 0.0f;
    for (int i = 0; i < M; ++i) tot += (key[i] < p) ? key[i] : 0.0f;       // :236-237
    s_tot = tot;
  }
  __syncthreads();
  const float p = s_p;
  for (int i = tid; i < M; i += NT) rw[i] = ((key[i] < p) ? key[i] : 0.0f) / s_tot;
  __syncthreads();
  cta_exclusive_scan<AGSF_WARPS>(rw, E, M, scratch);
  const float total = E[M];
  const int levels = bisect_levels(M);
  for (int i = tid; i < M; i += NT) ridx[i] = bisect_left_exclusive(E, M, levels, __fmul_rn(total, __fsub_rn(1.0f, u[i])));   // :238
  __syncthreads();
  // final weights of the last N entries, normalised (:241-244)
  if (tid == 0) {
    float t = 0.0f;
    for (int i = M - N; i < M; ++i) t += (key[i] < p) ? p : key[i];
    s_tot = t;
  }
  __syncthreads();
  for (int i = tid; i < N; i += NT) {
    const int s = M - N + i;
    const bool small = key[s] < p;
    idx_out[i] = small ? id[ridx[s]] : id[s];
    w_out[i] = (small ? p : key[s]) / s_tot;
  }
}
#endif  // BF_TU_MAIN
