// gsf_large.cuh -- Gaussian-sum (EKF) filter for LARGE states: Lorenz-96, dx up to 64 (BASELINE configs[3]: dx = 40,
// dy = 20, M = 256 components x 4096 trials).
//
// Same recursion as gsf_kernels.cuh (gaussfiltax/inference.py:333-365: `_condition_on` :72-105 on the carried
// prediction, max-shifted reweighting :347-350, `_predict` :51-70), but a component's covariance is 6.4 KB: it cannot
// live in registers, and the M components of a trial (1.6 MB) do not fit one SM's shared memory, so the carry
// (predicted mean and covariance of every component) STREAMS through HBM once per step -- read, update, predict,
// write back, 8 (dx + dx^2) bytes per component-step -- while one component at a time is worked on in a warp's
// shared-memory workspace.  One CTA per trial; its warps take the trial's components in turn; the only coupling
// (max and sum of the M log-likelihoods) happens in shared memory after the pass.  Predict does not depend on the
// weights, so update and predict of a component are fused in one visit.
//
// Model (oracle/models.py L96; gaussfiltax/nonlinearities.py:37-52 with the advection term restored):
//   f_i = x_i + dt ((x_{i+1} - x_{i-2}) x_{i-1} - x_i + F) + q_i         Jacobian: 4 non-zeros per row (cyclic band)
//   h_j = c1 x_{2j} + c2 x_{2j}^2 + r_j,  j < dx/2                       Jacobian: 1 non-zero per row
// Q and R must be diagonal here (they are in the configs).  Arithmetic mirrors the reference: K = solve(S + 1e-6
// [every element], H P)^T by LU with partial pivoting, P+ = P - (K S) K^T with the un-jittered S, log-likelihood by
// Cholesky of the un-jittered S.
#pragma once
#include "common.cuh"
#include "../../include/bfilt.h"

#define GL_MAXDX 64
#define GL_WARPS 6                  // 104 KB of shared memory per CTA at dx = 40: two CTAs (12 warps) per SM

struct LargeConst {                 // lives in global memory (too big for a kernel parameter)
  int dx, dy;
  float F, dt, c1, c2;
  float q0[GL_MAXDX], qd[GL_MAXDX], m0[GL_MAXDX];       // dynamics noise bias / variance (diagonal Q), initial mean
  float r0[GL_MAXDX / 2], rd[GL_MAXDX / 2];             // emission noise bias / variance (diagonal R)
  alignas(16) float P0[GL_MAXDX * GL_MAXDX];       // read with 128-bit loads
};

struct GsfLargeArgs {
  int M, T, resume, emissions_per_trial, flags;
  const LargeConst* lc;
  const float* emissions; const float* inputs;
  const float* init_means; const float* init_covs; const float* init_weights;
  float* carry_w; float* carry_m; float* carry_P;        // [B,M], [B,M,dx], [B,M,dx,dx]  (predicted moments)
  float* post_m;                                         // [B,M,dx] scratch: posterior means of the current step
  float* out_weights; float* out_means; float* out_covs; float* out_pred_means; float* out_pred_covs;
  float* out_loglik; float* out_mix_mean;
  const float* Q_t; const float* R_t;                    // [T,dx,dx] / [T,dy,dy] time-varying (diagonal) covariances or NULL
};

static inline size_t gsf_large_smem_bytes(int dx, int dy, int M) {
  // rounded to 4 floats: every warp's workspace (its P is read with 128-bit accesses) stays 16-byte aligned for any even dx
  const size_t per_warp = ((size_t)dx * dx + 2 * (size_t)dy * dx + 2 * (size_t)dy * dy + 4 * (size_t)dx + 4 * (size_t)dy + 8 + 3) & ~(size_t)3;
  return (GL_WARPS * per_warp + (size_t)M + 64) * sizeof(float);
}

// DXT: compile-time state dimension (0 = read it from LargeConst): index arithmetic folds, and for DXT = 40 the two
// rank-dy products are register-tiled.
template <int DXT>
__global__ void __launch_bounds__(GL_WARPS * 32, 2)
k_gsf_large(const __grid_constant__ GsfLargeArgs a) {
  extern __shared__ __align__(16) float smem[];
  const LargeConst& lc = *a.lc;
  const int dx = DXT ? DXT : lc.dx, dy = DXT ? DXT / 2 : lc.dy, M = a.M, T = a.T, b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t per_warp = ((size_t)dx * dx + 2 * (size_t)dy * dx + 2 * (size_t)dy * dy + 4 * (size_t)dx + 4 * (size_t)dy + 8 + 3) & ~(size_t)3;
  float* ws = smem + warp * per_warp;
  float* P = ws;                         // [dx,dx]
  float* X = P + dx * dx;                // [dy,dx]  H P, then X = solve(S', H P)  (K = X^T)
  float* KS = X + dy * dx;               // [dx,dy]  K S ; X and KS together are reused as F P [dx,dx] in the predict
  float* S = KS + dx * dy;               // [dy,dy]  un-jittered innovation covariance
  float* Sj = S + dy * dy;               // [dy,dy]  S + 1e-6, LU-factored in place; later the Cholesky factor of S
  float* mv = Sj + dy * dy;              // [dx] mean
  float* mn = mv + dx;                   // [dx] new mean
  float* fa = mn + dx;                   // [dx] dt * x_{i-1}
  float* fb = fa + dx;                   // [dx] dt * (x_{i+1} - x_{i-2})
  float* hv = fb + dx;                   // [dy] dh_j/dx_{2j}
  float* dv = hv + dy;                   // [dy] innovation
  float* zv = dv + dy;                   // [dy] scratch
  float* lr = zv + dy;                   // [dy] LU multipliers of the current column
  float* s_ll = smem + GL_WARPS * per_warp;      // [M] log-likelihoods, then normalised weights
  float* s_red = s_ll + M;                        // [64]
  const float jitter = (a.flags & BF_GSF_NO_JITTER) ? 0.0f : 1e-6f;
  const bool cov_plus_p = (a.flags & BF_GSF_COV_PLUS_P) != 0;
  const float dt = lc.dt, onemdt = 1.0f - lc.dt;

  auto wrap = [&](int i) { return i < 0 ? i + dx : (i >= dx ? i - dx : i); };
  const float* qdp = lc.qd; int qds = 1;               // diagonal of Q / R of the current step: element i at [i * stride]
  const float* rdp = lc.rd; int rds = 1;

  // ---- EKF predict of the component in (mv, P): mv <- f(mv), P <- F P F^T + Q     (inference.py:51-70)
  auto predict = [&](float u) {
    for (int i = lane; i < dx; i += 32) {
      const float xm1 = mv[wrap(i - 1)], xm2 = mv[wrap(i - 2)], xp1 = mv[wrap(i + 1)];
      fa[i] = dt * xm1;
      fb[i] = dt * (xp1 - xm2);
      mn[i] = mv[i] + dt * ((xp1 - xm2) * xm1 - mv[i] + lc.F) + lc.q0[i];
    }
    __syncwarp();
    float* FP = X;                       // [dx,dx] over X and KS
    // row i of F: (i-2): -fa_i, (i-1): fb_i, (i): 1-dt, (i+1): fa_i   (oracle/models.py L96.Fx, summed in column order)
    for (int idx = lane; idx < dx * dx; idx += 32) {
      const int i = idx / dx, c = idx - i * dx;
      const int j2 = wrap(i - 2), j1 = wrap(i - 1), jp = wrap(i + 1);
      FP[idx] = -fa[i] * P[j2 * dx + c] + fb[i] * P[j1 * dx + c] + onemdt * P[i * dx + c] + fa[i] * P[jp * dx + c];
    }
    __syncwarp();
    for (int idx = lane; idx < dx * dx; idx += 32) {
      const int i = idx / dx, c = idx - i * dx;
      const int j2 = wrap(c - 2), j1 = wrap(c - 1), jp = wrap(c + 1);
      const float v = -fa[c] * FP[i * dx + j2] + fb[c] * FP[i * dx + j1] + onemdt * FP[i * dx + c] + fa[c] * FP[i * dx + jp];
      const float add = cov_plus_p ? P[idx] : ((i == c) ? qdp[i * qds] : 0.0f);     // this loop reads FP only, P[idx] is this lane's own
      P[idx] = v + add;
    }
    for (int i = lane; i < dx; i += 32) mv[i] = mn[i];
    __syncwarp();
  };

  for (int t = 0; t < T; ++t) {
    if (a.Q_t) { qdp = a.Q_t + (size_t)t * dx * dx; qds = dx + 1; }       // `_get_params` (inference.py:21): step t's matrices
    if (a.R_t) { rdp = a.R_t + (size_t)t * dy * dy; rds = dy + 1; }
    const float* y = a.emissions + ((size_t)(a.emissions_per_trial ? b : 0) * T + t) * dy;
    const float u = a.inputs ? a.inputs[t] : 0.0f;
    const size_t bt = (size_t)b * T + t;

    for (int m = warp; m < M; m += GL_WARPS) {
      const size_t bm = (size_t)b * M + m;
      // ---- load the carried prediction
      const bool fresh = (t == 0 && !a.resume);
      const float* msrc = fresh ? a.init_means + bm * dx : a.carry_m + bm * dx;
      const float* Psrc = fresh ? (a.init_covs ? a.init_covs + bm * dx * dx : lc.P0) : a.carry_P + bm * dx * dx;
      for (int i = lane; i < dx; i += 32) mv[i] = msrc[i];
      if (!fresh && m + GL_WARPS < M) {                        // this warp's NEXT component: start its 6.4 KB on the way to L2 now
        const char* nxt = reinterpret_cast<const char*>(a.carry_P + (bm + GL_WARPS) * dx * dx);
        for (int off = lane * 128; off < dx * dx * 4; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + off));
      }
#pragma unroll 4
      for (int i = lane; i < dx * dx / 4; i += 32)           // dx is even: dx^2 is a multiple of 4, rows are 16-byte aligned
        reinterpret_cast<float4*>(P)[i] = reinterpret_cast<const float4*>(Psrc)[i];
      __syncwarp();
      if (fresh && (a.flags & BF_GSF_PREDICT_FIRST)) predict(u);

      // ---- `_condition_on`  (inference.py:72-105)
      for (int j = lane; j < dy; j += 32) {
        const float xs = mv[2 * j];
        hv[j] = lc.c1 + 2.0f * lc.c2 * xs;
        dv[j] = y[j] - (lc.c1 * xs + lc.c2 * xs * xs + lc.r0[j]);
      }
      __syncwarp();
      if constexpr (DXT == 40) {
        // ---- K = solve(S + 1e-6, H P)^T in REGISTERS.  A lane owns columns `lane` and `lane + 32` of the 20 x 60 matrix
        // [S' | H P] (columns 0-19: S', 20-59: the 40 right-hand sides), 20 rows each.  S' = S + 1e-6 11^T is symmetric
        // positive definite, so Gaussian elimination needs no pivoting (growth factor 1; jnp.linalg.solve's partial
        // pivoting changes the rounding, not the conditioning).  Multipliers and pivots travel by warp shuffle; every index
        // is a compile-time constant after unrolling.  This replaces a shared-memory LU (4 instructions per FMA) that was
        // half of the kernel's instructions.
        float c0[20], c1[20];
        const bool sc = lane < 20;                       // slot 0 holds an S' column (else the right-hand side lane - 20)
        const int col0 = sc ? 2 * lane : lane - 20;      // column of P behind slot 0
        const int col1 = (lane < 28) ? lane + 12 : 0;    // slot 1: right-hand side lane + 12 (lanes 28-31: idle copy)
        const float hl = sc ? hv[lane] : 1.0f;
        const float rl = sc ? rdp[lane * rds] : 0.0f;
#pragma unroll
        for (int j = 0; j < 20; ++j) {
          const float hj = hv[j];
          float v0 = hj * P[(2 * j) * 40 + col0];          // (H P)[j][col0]
          c1[j] = hj * P[(2 * j) * 40 + col1];
          if (sc) {
            v0 = ((j == lane) ? rl : 0.0f) + v0 * hl;      // S = H_r R H_r^T + H P H^T   (:99)
            S[j * 20 + lane] = v0;
            v0 += jitter;                                  // psd_solve: + 1e-6 on EVERY element
          }
          c0[j] = v0;
        }
#pragma unroll
        for (int k = 0; k < 20; ++k) {
          const float rp = __frcp_rn(__shfl_sync(BF_FULL, c0[k], k));
#pragma unroll
          for (int r = k + 1; r < 20; ++r) {
            const float l = __shfl_sync(BF_FULL, c0[r], k) * rp;
            c0[r] = fmaf(-l, c0[k], c0[r]);
            c1[r] = fmaf(-l, c1[k], c1[r]);
          }
        }
#pragma unroll
        for (int k = 19; k >= 0; --k) {                    // back substitution on the right-hand-side slots only
          const float rp = __frcp_rn(__shfl_sync(BF_FULL, c0[k], k));
          float a0 = c0[k], a1 = c1[k];
#pragma unroll
          for (int j = k + 1; j < 20; ++j) {
            const float uu = __shfl_sync(BF_FULL, c0[k], j);      // U[k][j]
            a0 = fmaf(-uu, c0[j], a0);
            a1 = fmaf(-uu, c1[j], a1);
          }
          if (!sc) c0[k] = a0 * rp;
          c1[k] = a1 * rp;
        }
#pragma unroll
        for (int k = 0; k < 20; ++k) {                     // X = solve(S', H P)  (K = X^T) back to shared memory for the tiled products
          if (!sc) X[k * 40 + lane - 20] = c0[k];
          if (lane < 28) X[k * 40 + lane + 12] = c1[k];
        }
        __syncwarp();
      } else {
      for (int idx = lane; idx < dy * dx; idx += 32) {           // H P
          const int j = idx / dx, c = idx - j * dx;
          X[idx] = hv[j] * P[(2 * j) * dx + c];
        }
        __syncwarp();
        for (int idx = lane; idx < dy * dy; idx += 32) {           // S = H_r R H_r^T + H P H^T ; S' = S + 1e-6
          const int j = idx / dy, k = idx - j * dy;
          const float s = ((j == k) ? rdp[j * rds] : 0.0f) + X[j * dx + 2 * k] * hv[k];
          S[idx] = s; Sj[idx] = s + jitter;
        }
        __syncwarp();
        // LU with partial pivoting of S', applied to the dx right-hand sides in X   (jnp.linalg.solve, utils.py:259)
        for (int k = 0; k < dy; ++k) {
          float best = -1.0f; int bi = k;
          for (int r = k + lane; r < dy; r += 32) { const float v = fabsf(Sj[r * dy + k]); if (v > best) { best = v; bi = r; } }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(BF_FULL, best, o); const int oi = __shfl_xor_sync(BF_FULL, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
          }
          if (bi != k) {
            for (int c = lane; c < dy; c += 32) { const float t0 = Sj[k * dy + c]; Sj[k * dy + c] = Sj[bi * dy + c]; Sj[bi * dy + c] = t0; }
            for (int c = lane; c < dx; c += 32) { const float t0 = X[k * dx + c]; X[k * dx + c] = X[bi * dx + c]; X[bi * dx + c] = t0; }
          }
          __syncwarp();
          const float piv = Sj[k * dy + k];
          for (int r = k + 1 + lane; r < dy; r += 32) lr[r] = Sj[r * dy + k] / piv;
          __syncwarp();
          // rank-1 update of the trailing block of S' and of all of X: a lane owns COLUMNS of the concatenation
          // [S'(:, k+1:) | X] and walks down the rows -- no index division in the loop (it was 38 % of the kernel's
          // instructions when the block was flattened over (row, column)), the pivot-row element stays in a register
          const int ncs = dy - k - 1;
          for (int cc = lane; cc < ncs + dx; cc += 32) {
            float* col = (cc < ncs) ? (Sj + k + 1 + cc) : (X + (cc - ncs));
            const int stride = (cc < ncs) ? dy : dx;
            const float top = col[k * stride];
            for (int r = k + 1; r < dy; ++r) col[r * stride] -= lr[r] * top;
          }
          __syncwarp();
        }
        for (int k = dy - 1; k >= 0; --k) {                        // back substitution, one column of X per lane
          for (int c = lane; c < dx; c += 32) {
            float s = X[k * dx + c];
            for (int j = k + 1; j < dy; ++j) s -= Sj[k * dy + j] * X[j * dx + c];
            X[k * dx + c] = s / Sj[k * dy + k];
          }
          __syncwarp();
        }
      }
      float pt[DXT == 40 ? 5 : 1][DXT == 40 ? 10 : 1];     // DXT = 40: this lane's 5 x 10 tile of the posterior covariance
      if constexpr (DXT == 40) {
        // register tiles: lane = (row group of 5) x (column group): 25 / 50 accumulators, 10 / 15 loads per 25 / 50 FMAs
        const int rg = lane >> 2, cg = lane & 3;
        {
          float acc[5][5];
#pragma unroll
          for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc[i][j] = 0.0f;
          for (int k = 0; k < 20; ++k) {                           // K S,  K = X^T
            float xk[5], sk[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) xk[i] = X[k * 40 + 5 * rg + i];
#pragma unroll
            for (int j = 0; j < 5; ++j) sk[j] = S[k * 20 + 5 * cg + j];
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
              for (int j = 0; j < 5; ++j) acc[i][j] += xk[i] * sk[j];
          }
#pragma unroll
          for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) KS[(5 * rg + i) * 20 + 5 * cg + j] = acc[i][j];
        }
        __syncwarp();
        {
          float acc[5][10];
#pragma unroll
          for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 10; ++j) acc[i][j] = 0.0f;
          for (int k = 0; k < 20; ++k) {                           // (K S) K^T
            float ks[5], xr[10];
#pragma unroll
            for (int i = 0; i < 5; ++i) ks[i] = KS[(5 * rg + i) * 20 + k];
#pragma unroll
            for (int j = 0; j < 10; ++j) xr[j] = X[k * 40 + 10 * cg + j];
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
              for (int j = 0; j < 10; ++j) acc[i][j] += ks[i] * xr[j];
          }
#pragma unroll
          for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 10; ++j) pt[i][j] = P[(5 * rg + i) * 40 + 10 * cg + j] - acc[i][j];      // P - (K S) K^T  (inference.py:101)
          if (a.out_covs) {                                  // only a requested output needs the posterior back in shared memory
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
              for (int j = 0; j < 10; ++j) P[(5 * rg + i) * 40 + 10 * cg + j] = pt[i][j];
          }
        }
      } else {
        for (int idx = lane; idx < dx * dy; idx += 32) {           // K S,  K = X^T
          const int i = idx / dy, j = idx - i * dy;
          float s = 0.0f;
          for (int k = 0; k < dy; ++k) s += X[k * dx + i] * S[k * dy + j];
          KS[idx] = s;
        }
        __syncwarp();
        for (int idx = lane; idx < dx * dx; idx += 32) {           // P - (K S) K^T      (inference.py:101)
          const int i = idx / dx, c = idx - i * dx;
          float s = 0.0f;
          for (int k = 0; k < dy; ++k) s += KS[i * dy + k] * X[k * dx + c];
          P[idx] -= s;
        }
      }
      __syncwarp();
      for (int i = lane; i < dx; i += 32) {                      // m + K (y - h(m))   (inference.py:102)
        float s = 0.0f;
        for (int k = 0; k < dy; ++k) s += X[k * dx + i] * dv[k];
        mv[i] += s;
      }
      __syncwarp();
      // log N(y; h(m), S) with the un-jittered S (TFP: Cholesky + forward substitution)
      float logdet = 0.0f, ss = 0.0f;
      if constexpr (DXT == 40) {
        // S = L D L^T by the same register-resident elimination (lane k < 20 owns column k of S, lane 20 the innovation):
        // z' = L^-1 d,  ||chol(S)^-1 d||^2 = sum z'_k^2 / D_k,  sum log diag chol(S) = 0.5 sum log D_k
        float c0[20];
#pragma unroll
        for (int j = 0; j < 20; ++j) c0[j] = (lane < 20) ? S[j * 20 + lane] : ((lane == 20) ? dv[j] : 0.0f);
#pragma unroll
        for (int k = 0; k < 20; ++k) {
          const float piv = __shfl_sync(BF_FULL, c0[k], k);
          const float zk = __shfl_sync(BF_FULL, c0[k], 20);
          const float rp = __frcp_rn(piv);
          logdet += 0.5f * logf(piv);
          ss = fmaf(zk * zk, rp, ss);
#pragma unroll
          for (int r = k + 1; r < 20; ++r) {
            const float l = __shfl_sync(BF_FULL, c0[r], k) * rp;
            c0[r] = fmaf(-l, c0[k], c0[r]);
          }
        }
      } else {
      // log N(y; h(m), S): Cholesky of the un-jittered S (TFP), column by column, in Sj
        for (int j = 0; j < dy; ++j) {                             // dy <= 32: lane l owns row j + l of column j
          const int i = j + lane;
          float v = 0.0f;
          if (i < dy) {
            v = S[i * dy + j];
            for (int k = 0; k < j; ++k) v -= Sj[i * dy + k] * Sj[j * dy + k];
          }
          const float ljj = sqrtf(__shfl_sync(BF_FULL, v, 0));     // lane 0 holds the diagonal element
          logdet += logf(ljj);
          if (i < dy) Sj[i * dy + j] = (i == j) ? ljj : v / ljj;
          __syncwarp();
        }
        for (int i = 0; i < dy; ++i) {                             // forward substitution (uniform across the warp)
          float s = dv[i];
          for (int k = 0; k < i; ++k) s -= Sj[i * dy + k] * zv[k];
          const float z = s / Sj[i * dy + i];
          __syncwarp();
          if (lane == 0) zv[i] = z;
          __syncwarp();
          ss += z * z;
        }
      }
      if (lane == 0) s_ll[m] = -0.5f * ss - logdet - 0.5f * (float)dy * 1.8378770664093453f;
      // posterior outputs, and the posterior mean for the mixture mean after the reweighting
      const size_t btm = bt * M + m;
      for (int i = lane; i < dx; i += 32) {
        a.post_m[bm * dx + i] = mv[i];
        if (a.out_means) a.out_means[btm * dx + i] = mv[i];
      }
      if (a.out_covs) for (int i = lane; i < dx * dx; i += 32) a.out_covs[btm * dx * dx + i] = P[i];
      __syncwarp();
      // ---- `_predict`, then the carry goes back to HBM
      if constexpr (DXT == 40) {
        // F P F^T on the 5 x 10 register tiles: F has 4 non-zeros per row (columns i-2, i-1, i, i+1, cyclic), so F P mixes a
        // tile with two rows of the lane above and one of the lane below, (F P) F^T with two columns of the lane to the left
        // and one to the right -- warp shuffles; the covariance never returns to shared memory and goes to HBM from registers
        const int rg = lane >> 2, cg = lane & 3;
        for (int i = lane; i < 40; i += 32) {
          const float xm1 = mv[wrap(i - 1)], xm2 = mv[wrap(i - 2)], xp1 = mv[wrap(i + 1)];
          fa[i] = dt * xm1;
          fb[i] = dt * (xp1 - xm2);
          mn[i] = mv[i] + dt * ((xp1 - xm2) * xm1 - mv[i] + lc.F) + lc.q0[i];
        }
        __syncwarp();
        const int up = (((rg + 7) & 7) << 2) | cg, dn = (((rg + 1) & 7) << 2) | cg;
        const int lf = (rg << 2) | ((cg + 3) & 3), rt = (rg << 2) | ((cg + 1) & 3);
        float fp[5][10];
        {
          float e0[10], e1[10], e7[10];                    // rows i-2, i-1 of the first tile row; row i+1 of the last
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            e0[j] = __shfl_sync(BF_FULL, pt[3][j], up); e1[j] = __shfl_sync(BF_FULL, pt[4][j], up);
            e7[j] = __shfl_sync(BF_FULL, pt[0][j], dn);
          }
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const float far = fa[5 * rg + r], fbr = fb[5 * rg + r];
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              const float m2 = (r == 0) ? e0[j] : (r == 1) ? e1[j] : pt[r >= 2 ? r - 2 : 0][j];
              const float m1 = (r == 0) ? e1[j] : pt[r >= 1 ? r - 1 : 0][j];
              const float p1 = (r == 4) ? e7[j] : pt[r < 4 ? r + 1 : 4][j];
              fp[r][j] = -far * m2 + fbr * m1 + onemdt * pt[r][j] + far * p1;
            }
          }
        }
        float fac[10], fbc[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) { fac[j] = fa[10 * cg + j]; fbc[j] = fb[10 * cg + j]; }
        float* cP = a.carry_P + bm * 1600;
        float* oP = a.out_pred_covs ? a.out_pred_covs + btm * 1600 : nullptr;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const float l0 = __shfl_sync(BF_FULL, fp[r][8], lf), l1 = __shfl_sync(BF_FULL, fp[r][9], lf);
          const float r0 = __shfl_sync(BF_FULL, fp[r][0], rt);
          float o[10];
#pragma unroll
          for (int j = 0; j < 10; ++j) {
            const float m2 = (j == 0) ? l0 : (j == 1) ? l1 : fp[r][j >= 2 ? j - 2 : 0];
            const float m1 = (j == 0) ? l1 : fp[r][j >= 1 ? j - 1 : 0];
            const float p1 = (j == 9) ? r0 : fp[r][j < 9 ? j + 1 : 9];
            const float v = -fac[j] * m2 + fbc[j] * m1 + onemdt * fp[r][j] + fac[j] * p1;
            const int gi = 5 * rg + r, gc = 10 * cg + j;
            o[j] = v + (cov_plus_p ? pt[r][j] : ((gi == gc) ? qdp[gi * qds] : 0.0f));
          }
          float2* dst = reinterpret_cast<float2*>(cP + (5 * rg + r) * 40 + 10 * cg);       // 40-byte aligned rows of 10 floats
#pragma unroll
          for (int j = 0; j < 5; ++j) dst[j] = make_float2(o[2 * j], o[2 * j + 1]);
          if (oP) {
            float2* d2 = reinterpret_cast<float2*>(oP + (5 * rg + r) * 40 + 10 * cg);
#pragma unroll
            for (int j = 0; j < 5; ++j) d2[j] = make_float2(o[2 * j], o[2 * j + 1]);
          }
        }
        for (int i = lane; i < 40; i += 32) {
          a.carry_m[bm * 40 + i] = mn[i];
          if (a.out_pred_means) a.out_pred_means[btm * 40 + i] = mn[i];
        }
      } else {
        predict(u);
        for (int i = lane; i < dx; i += 32) {
          a.carry_m[bm * dx + i] = mv[i];
          if (a.out_pred_means) a.out_pred_means[btm * dx + i] = mv[i];
        }
#pragma unroll 4
        for (int i = lane; i < dx * dx / 4; i += 32)
          reinterpret_cast<float4*>(a.carry_P + bm * dx * dx)[i] = reinterpret_cast<const float4*>(P)[i];
        if (a.out_pred_covs) for (int i = lane; i < dx * dx; i += 32) {
          a.out_pred_covs[btm * dx * dx + i] = P[i];
        }
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- lls -= max ; w = exp(lls) * w ; w /= sum(w)       (inference.py:347-350)
    float mx = -INFINITY;
    for (int m = tid; m < M; m += GL_WARPS * 32) mx = fmaxf(mx, s_ll[m]);
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int w = 1; w < GL_WARPS; ++w) mx = fmaxf(mx, s_red[w]);
    float z = 0.0f;
    for (int m = tid; m < M; m += GL_WARPS * 32) {
      const size_t bm = (size_t)b * M + m;
      const float wprev = (t == 0 && !a.resume) ? (a.init_weights ? a.init_weights[bm] : 1.0f / (float)M) : a.carry_w[bm];
      const float e = expf(s_ll[m] - mx) * wprev;
      s_ll[m] = e; z += e;
    }
    z = warp_sum(z);
    __syncthreads();
    if (lane == 0) s_red[32 + warp] = z;
    __syncthreads();
    float Z = 0.0f;
    for (int w = 0; w < GL_WARPS; ++w) Z += s_red[32 + w];
    for (int m = tid; m < M; m += GL_WARPS * 32) {
      const float w = s_ll[m] / Z;
      s_ll[m] = w;
      a.carry_w[(size_t)b * M + m] = w;
      if (a.out_weights) a.out_weights[bt * M + m] = w;
    }
    if (tid == 0 && a.out_loglik) a.out_loglik[bt] = mx + logf(Z);
    __syncthreads();
    if (a.out_mix_mean) {
      for (int i = tid; i < dx; i += GL_WARPS * 32) {
        float s = 0.0f;
        for (int m = 0; m < M; ++m) s += s_ll[m] * a.post_m[((size_t)b * M + m) * dx + i];
        a.out_mix_mean[bt * dx + i] = s;
      }
    }
    __syncthreads();
  }
}
