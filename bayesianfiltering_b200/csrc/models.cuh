// models.cuh -- compiled model registry: f(x,q,u), h(x,r,u) and their analytic Jacobians as
// __device__ functors templated into the filter kernels.  The reference takes arbitrary Python
// callables (gaussfiltax/models.py:46-50) and differentiates them with jacfwd (inference.py:327-329);
// a kernel cannot, so the models its drivers define are compiled in.  Sources of each definition are
// listed in oracle/models.py, which restates the same functions in NumPy and is what these are tested
// against (values at rtol 1e-5, Jacobians also against float64 finite differences).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// UNGM / "stochastic growth model":  f = a x + b x/(1+x^2) + c u + q ;  h = d x + e x^2 + r
// theta = [a b c d e]     (Experiment_TSP_2023.ipynb cell 2: f3, g3)
struct UngmModel {
  static constexpr int DX = 1, DQ = 1, DY = 1, DR = 1;
  __device__ __forceinline__ static void f(const float* x, const float* q, float u, const ModelConst& c, float* o) {
    const float xx = x[0];
    const float inv = __frcp_rn(1.0f + xx * xx);          // correctly rounded reciprocal, shared with Fx below
    o[0] = fmaf(c.theta[2], u, fmaf(c.theta[1] * xx, inv, c.theta[0] * xx)) + q[0];
  }
  __device__ __forceinline__ static void h(const float* x, const float* r, float u, const ModelConst& c, float* o) {
    o[0] = fmaf(c.theta[4] * x[0], x[0], c.theta[3] * x[0]) + r[0];
  }
  __device__ __forceinline__ static void Fx(const float* x, const float* q, float u, const ModelConst& c, float* J) {
    const float x2 = x[0] * x[0], inv = __frcp_rn(1.0f + x2);
    J[0] = c.theta[0] + c.theta[1] * (1.0f - x2) * inv * inv;
  }
  __device__ __forceinline__ static void Fq(const float* x, const float* q, float u, const ModelConst& c, float* J) { J[0] = 1.0f; }
  __device__ __forceinline__ static void Hx(const float* x, const float* r, float u, const ModelConst& c, float* J) {
    J[0] = c.theta[3] + 2.0f * c.theta[4] * x[0];
  }
};

// sin and cos to ~1 ulp: minimax polynomials (Cephes sinf/cosf kernels) for |x| <= pi/4, libm otherwise.
// (fma / mul spelled out here and below: the particle-filter translation units are compiled with -fmad=false, so that
// the general, lean and sharded instantiations of a kernel cannot round a product-sum differently.)
__device__ __forceinline__ void sincos_small(float x, float* sn, float* cs) {
  if (fabsf(x) <= 0.78539816f) {
    const float z = x * x;
    const float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    *sn = fmaf(x * z, ps, x);
    const float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    *cs = fmaf(z * z, pc, fmaf(-0.5f, z, 1.0f));
  } else {
    sincosf(x, sn, cs);
  }
}

// atan2 to ~1 ulp in ~35 instructions (libm's atan2f is ~52): one division, Cephes atanf's reduction at
// tan(pi/8) and its degree-4 polynomial in t^2, then the quadrant.
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const bool mid = mn > 0.41421356f * mx;               // atan(mn/mx) = pi/4 + atan((mn-mx)/(mn+mx))
  const float num = mid ? mn - mx : mn, den = mid ? mn + mx : mx;
  const float t = (den == 0.0f) ? 0.0f : __fdividef(num, den);
  const float z = t * t;
  const float p = fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
  float r = fmaf(p * z, t, t);
  if (mid) r = (r + -2.18556941e-8f) + 0.785398185253143310546875f;     // pi/4 as hi + lo
  if (ay > ax) r = 1.57079632679489662f - r;
  if (x < 0.0f) r = 3.14159265358979324f - r;
  return copysignf(r, y);
}

// ---------------------------------------------------------------------------------------------
// Manoeuvring bearings-only tracking (BOT_Experiment_script.py:27-47).  theta = [dt acc omega_scale]
template <int DY_>
struct BotModel {
  static constexpr int DX = 4, DQ = 2, DY = DY_, DR = DY_;

  struct Turn { float s, om, sn, cs, A, B; };
  __device__ __forceinline__ static Turn turn(const float* x, const ModelConst& c) {
    Turn t;
    // one rsqrt instead of sqrt + three divisions (each result within 2 ulp of the reference's expression)
    const float v2 = fmaf(x[1], x[1], x[3] * x[3]);
    const float rs = rsqrtf(v2);
    const float k = c.theta[2] * c.theta[1];
    t.s = v2 * rs;
    t.om = k * rs;
    sincos_small(c.theta[0] * t.om, &t.sn, &t.cs);
    const float iom = t.s * __frcp_rn(k);               // k is uniform: the reciprocal is hoisted
    t.A = t.sn * iom;
    t.B = (1.0f - t.cs) * iom;
    return t;
  }
  __device__ __forceinline__ static void coef(float u, float& c0, float& c1, float& c2) {
    c0 = 0.5f * (u - 1.0f) * (u - 2.0f); c1 = -u * (u - 2.0f); c2 = 0.5f * u * (u - 1.0f);
  }
  // (c0 FCV + c1 FCT(x, acc) + c2 FCT(x, -acc)) x + G q, accumulated as fma(c2, ., fma(c1, ., fma(c0, ., G q))).  The
  // manoeuvre input u is the same for every particle of a step, and for u in {0, 1, 2} two of the three coefficients are
  // exactly 0: a term whose coefficient is 0 contributes exactly 0 and is skipped (a uniform branch) -- unless the speed is
  // 0, where the reference's 0 * NaN poisons the result and so does this.
  __device__ __forceinline__ static void f(const float* x, const float* q, float u, const ModelConst& c, float* o) {
    float c0, c1, c2; coef(u, c0, c1, c2);
    const float T = c.theta[0];
    float r0 = 0.5f * q[0], r1 = q[0], r2 = 0.5f * q[1], r3 = q[1];
    const bool still = (x[1] == 0.0f && x[3] == 0.0f);
    if (c0 != 0.0f) {
      r0 = fmaf(c0, fmaf(T, x[1], x[0]), r0); r1 = fmaf(c0, x[1], r1);
      r2 = fmaf(c0, fmaf(T, x[3], x[2]), r2); r3 = fmaf(c0, x[3], r3);
    }
    if (c1 != 0.0f || c2 != 0.0f || still) {
      const Turn t = turn(x, c);
      // FCT(x,+acc) x and FCT(x,-acc) x share sin/cos: Omega -> -Omega flips sn and B only
      const float p0 = fmaf(t.A, x[1], x[0]), p2 = fmaf(t.A, x[3], x[2]);
      const float cc1 = t.cs * x[1], cc3 = t.cs * x[3], ss1 = t.sn * x[1], ss3 = t.sn * x[3];
      if (c1 != 0.0f || still) {
        r0 = fmaf(c1, fmaf(-t.B, x[3], p0), r0); r1 = fmaf(c1, cc1 - ss3, r1);
        r2 = fmaf(c1, fmaf(t.B, x[1], p2), r2); r3 = fmaf(c1, ss1 + cc3, r3);
      }
      if (c2 != 0.0f || still) {
        r0 = fmaf(c2, fmaf(t.B, x[3], p0), r0); r1 = fmaf(c2, cc1 + ss3, r1);
        r2 = fmaf(c2, fmaf(-t.B, x[1], p2), r2); r3 = fmaf(c2, cc3 - ss1, r3);
      }
    }
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
  }
  __device__ __forceinline__ static void h(const float* x, const float* r, float u, const ModelConst& c, float* o) {
    o[0] = atan2_fast(x[2], x[0]) + r[0];
    if (DY == 2) o[1] = sqrtf(fmaf(x[0], x[0], x[2] * x[2])) + r[1];
  }
  // d(FCT(x,a) x)/dx for a = +acc (sgn=+1) or -acc (sgn=-1), accumulated into J with weight cw
  __device__ __forceinline__ static void fct_jac(const float* x, const Turn& t, float T, float sgn, float cw, float* J) {
    const float om = sgn * t.om, sn = sgn * t.sn, cs = t.cs, A = t.A, B = sgn * t.B;
    const float dA = (T * cs * om - sn) / (om * om);
    const float dB = (T * sn * om - (1.0f - cs)) / (om * om);
    const float dsn = T * cs, dcs = -T * sn;
    const float v[4] = {dA * x[1] - dB * x[3], dcs * x[1] - dsn * x[3], dB * x[1] + dA * x[3], dsn * x[1] + dcs * x[3]};
    const float s2 = t.s * t.s;
    const float g1 = -om * x[1] / s2, g3 = -om * x[3] / s2;
    const float M[16] = {1.0f, A, 0.0f, -B, 0.0f, cs, 0.0f, -sn, 0.0f, B, 1.0f, A, 0.0f, sn, 0.0f, cs};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float e = M[i * 4 + j];
        if (j == 1) e += v[i] * g1;
        if (j == 3) e += v[i] * g3;
        J[i * 4 + j] += cw * e;
      }
    }
  }
  __device__ __forceinline__ static void Fx(const float* x, const float* q, float u, const ModelConst& c, float* J) {
    float c0, c1, c2; coef(u, c0, c1, c2);
    const float T = c.theta[0];
    const float CV[16] = {1.0f, T, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, T, 0.0f, 0.0f, 0.0f, 1.0f};
#pragma unroll
    for (int i = 0; i < 16; ++i) J[i] = c0 * CV[i];
    Turn t = turn(x, c);
    fct_jac(x, t, T, 1.0f, c1, J);
    fct_jac(x, t, T, -1.0f, c2, J);
  }
  __device__ __forceinline__ static void Fq(const float* x, const float* q, float u, const ModelConst& c, float* J) {
    const float G[8] = {0.5f, 0.0f, 1.0f, 0.0f, 0.0f, 0.5f, 0.0f, 1.0f};
#pragma unroll
    for (int i = 0; i < 8; ++i) J[i] = G[i];
  }
  __device__ __forceinline__ static void Hx(const float* x, const float* r, float u, const ModelConst& c, float* J) {
    const float d2 = x[0] * x[0] + x[2] * x[2];
    J[0] = -x[2] / d2; J[1] = 0.0f; J[2] = x[0] / d2; J[3] = 0.0f;
    if (DY == 2) {
      const float rg = sqrtf(d2);
      J[4] = x[0] / rg; J[5] = 0.0f; J[6] = x[2] / rg; J[7] = 0.0f;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Lorenz-63, Euler step; h = c x.x + r.  theta = [sigma rho beta dt c]
struct L63Model {
  static constexpr int DX = 3, DQ = 3, DY = 1, DR = 1;
  __device__ __forceinline__ static void f(const float* x, const float* q, float u, const ModelConst& c, float* o) {
    const float sg = c.theta[0], rho = c.theta[1], be = c.theta[2], dt = c.theta[3];
    o[0] = fmaf(dt * sg, x[1] - x[0], x[0]) + q[0];
    o[1] = fmaf(dt, fmaf(-x[0], x[2], fmaf(x[0], rho, -x[1])), x[1]) + q[1];
    o[2] = fmaf(dt, fmaf(x[0], x[1], -be * x[2]), x[2]) + q[2];
  }
  __device__ __forceinline__ static void h(const float* x, const float* r, float u, const ModelConst& c, float* o) {
    o[0] = fmaf(c.theta[4], fmaf(x[2], x[2], fmaf(x[1], x[1], x[0] * x[0])), r[0]);
  }
  __device__ __forceinline__ static void Fx(const float* x, const float* q, float u, const ModelConst& c, float* J) {
    const float sg = c.theta[0], rho = c.theta[1], be = c.theta[2], dt = c.theta[3];
    J[0] = 1.0f - dt * sg; J[1] = dt * sg; J[2] = 0.0f;
    J[3] = dt * (rho - x[2]); J[4] = 1.0f - dt; J[5] = -dt * x[0];
    J[6] = dt * x[1]; J[7] = dt * x[0]; J[8] = 1.0f - dt * be;
  }
  __device__ __forceinline__ static void Fq(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  }
  __device__ __forceinline__ static void Hx(const float* x, const float* r, float u, const ModelConst& c, float* J) {
    const float k = 2.0f * c.theta[4];
    J[0] = k * x[0]; J[1] = k * x[1]; J[2] = k * x[2];
  }
};

// ---------------------------------------------------------------------------------------------
// Linear-Gaussian: f = A x + q, h = H x + r.  theta = [A (DX*DX) | H (DY*DX)], row-major
template <int DX_, int DY_>
struct LgModel {
  static constexpr int DX = DX_, DQ = DX_, DY = DY_, DR = DY_;
  __device__ __forceinline__ static void f(const float* x, const float* q, float u, const ModelConst& c, float* o) {
#pragma unroll
    for (int i = 0; i < DX; ++i) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j < DX; ++j) a = fmaf(c.theta[i * DX + j], x[j], a);
      o[i] = a + q[i];
    }
  }
  __device__ __forceinline__ static void h(const float* x, const float* r, float u, const ModelConst& c, float* o) {
#pragma unroll
    for (int i = 0; i < DY; ++i) {
      float a = 0.0f;
#pragma unroll
      for (int j = 0; j < DX; ++j) a = fmaf(c.theta[DX * DX + i * DX + j], x[j], a);
      o[i] = a + r[i];
    }
  }
  __device__ __forceinline__ static void Fx(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) J[i] = c.theta[i];
  }
  __device__ __forceinline__ static void Fq(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) J[i] = (i % (DX + 1) == 0) ? 1.0f : 0.0f;
  }
  __device__ __forceinline__ static void Hx(const float* x, const float* r, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < DY * DX; ++i) J[i] = c.theta[DX * DX + i];
  }
};

// ---------------------------------------------------------------------------------------------
// ICASSP toy: f = sin(w x) + q (element-wise), h = c x.x + r.  theta = [w c]
template <int DX_>
struct Sin1dModel {
  static constexpr int DX = DX_, DQ = DX_, DY = 1, DR = 1;
  __device__ __forceinline__ static void f(const float* x, const float* q, float u, const ModelConst& c, float* o) {
#pragma unroll
    for (int i = 0; i < DX; ++i) o[i] = sinf(c.theta[0] * x[i]) + q[i];
  }
  __device__ __forceinline__ static void h(const float* x, const float* r, float u, const ModelConst& c, float* o) {
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < DX; ++i) a = fmaf(x[i], x[i], a);
    o[0] = fmaf(c.theta[1], a, r[0]);
  }
  __device__ __forceinline__ static void Fx(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) J[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < DX; ++i) J[i * DX + i] = c.theta[0] * cosf(c.theta[0] * x[i]);
  }
  __device__ __forceinline__ static void Fq(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < DX * DX; ++i) J[i] = (i % (DX + 1) == 0) ? 1.0f : 0.0f;
  }
  __device__ __forceinline__ static void Hx(const float* x, const float* r, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < DX; ++i) J[i] = 2.0f * c.theta[1] * x[i];
  }
};

// ---------------------------------------------------------------------------------------------
// Multivariate stochastic volatility with a switch (adaptive_experiment.py:47-57, fmsv / glmsv / lmsvlp):
//   f = Phi x + q (Phi diagonal);  h = u beta exp(x/sigma) * r + (1-u) (H0 x + r) (H0 diagonal): with u = 1 the emission
//   noise enters MULTIPLICATIVELY, so dh/dr = M(x,u) = diag(u beta exp(x/sigma) + (1-u)) is not the identity -- the
//   Gaussian-sum update uses S = M R M^T + H P H^T (inference.py:99) and the particle filter's log-prob has the
//   state-dependent covariance M R M^T.  theta = [Phi (D) | beta (D) | sigma | H0 (D)].
template <int D>
struct MsvModel {
  static constexpr int DX = D, DQ = D, DY = D, DR = D;
  static constexpr bool NONADDITIVE_R = true;
  __device__ __forceinline__ static void f(const float* x, const float* q, float u, const ModelConst& c, float* o) {
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = fmaf(c.theta[i], x[i], q[i]);
  }
  // diagonal of dh/dr at (x, u)
  __device__ __forceinline__ static void Hr_diag(const float* x, float u, const ModelConst& c, float* d) {
    const float isg = __frcp_rn(c.theta[2 * D]);
#pragma unroll
    for (int i = 0; i < D; ++i) d[i] = fmaf(u * c.theta[D + i], expf(x[i] * isg), 1.0f - u);
  }
  __device__ __forceinline__ static void h(const float* x, const float* r, float u, const ModelConst& c, float* o) {
    const float isg = __frcp_rn(c.theta[2 * D]);
#pragma unroll
    for (int i = 0; i < D; ++i)
      o[i] = fmaf(u * c.theta[D + i] * expf(x[i] * isg), r[i], (1.0f - u) * fmaf(c.theta[2 * D + 1 + i], x[i], r[i]));
  }
  __device__ __forceinline__ static void Fx(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) J[i] = (i % (D + 1) == 0) ? c.theta[i / (D + 1)] : 0.0f;
  }
  __device__ __forceinline__ static void Fq(const float* x, const float* q, float u, const ModelConst& c, float* J) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) J[i] = (i % (D + 1) == 0) ? 1.0f : 0.0f;
  }
  __device__ __forceinline__ static void Hx(const float* x, const float* r, float u, const ModelConst& c, float* J) {
    const float isg = __frcp_rn(c.theta[2 * D]);
#pragma unroll
    for (int i = 0; i < D * D; ++i) J[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i)
      J[i * D + i] = fmaf(u * c.theta[D + i] * expf(x[i] * isg) * isg, r[i], (1.0f - u) * c.theta[2 * D + 1 + i]);
  }
};

// does the model's emission noise enter non-additively (dh/dr != I)?
template <class M, class = void> struct nonadditive_r { static constexpr bool value = false; };
template <class M> struct nonadditive_r<M, decltype((void)M::NONADDITIVE_R)> { static constexpr bool value = M::NONADDITIVE_R; };

// ---------------------------------------------------------------------------------------------
// MVN(loc = h(x, r0, u), covariance_matrix = R).log_prob(y) -- the `*lp` callables of the reference's
// drivers (e.g. BOT_Experiment_script.py:45) with TFP's formula: z = L^-1 (y - loc) by forward
// substitution, -0.5 z.z - sum(log diag L) - 0.5 k log(2 pi).
template <class M>
__device__ __forceinline__ float emission_log_prob(const float* x, const float* y, float u, const ModelConst& c) {
  float loc[M::DY];
  M::h(x, c.r0, u, c, loc);
  float z[M::DY];
  float ss = 0.0f, lp = c.lp_const;
  float dd[M::DY];
  if constexpr (nonadditive_r<M>::value) {
    // covariance M R M^T with M = diag(d) > 0: its Cholesky factor is diag(d) L_R, so the whitened residual is
    // L_R^-1 diag(d)^-1 (y - loc) and log det gains sum(log d)   (adaptive_experiment.py:55-57 `lmsvlp`)
    M::Hr_diag(x, u, c, dd);
#pragma unroll
    for (int i = 0; i < M::DY; ++i) lp -= logf(dd[i]);
  }
#pragma unroll
  for (int i = 0; i < M::DY; ++i) {
    float d = y[i] - loc[i];
    if constexpr (nonadditive_r<M>::value) d = __fdiv_rn(d, dd[i]);
#pragma unroll
    for (int j = 0; j < i; ++j) d = fmaf(-c.LR[i * M::DR + j], z[j], d);
    z[i] = d * c.iLR[i];
    ss = fmaf(z[i], z[i], ss);
  }
  return fmaf(-0.5f, ss, lp);
}

// Dispatch a functor over the registry.  F must provide `template<class M> int run()`.  Returns -3 when the model is
// not instantiated HERE: the kernel translation units are compiled as BF_NGROUPS slices of the registry (entry k
// belongs to slice k % BF_NGROUPS), several nvcc processes side by side (csrc/build.sh).  Heavy entries come first so
// that they land in different slices.
#ifndef BF_NGROUPS
#define BF_NGROUPS 1
#define BF_GROUP 0
#endif
#ifdef BF_DEV_MODELS          // fast development builds: only the two bench models
#define BF_DEV_KEEP(k) ((k) == 1 || (k) == 2)
#else
#define BF_DEV_KEEP(k) true
#endif
#define BF_ENTRY(k, cond, ...)                                                        \
  if constexpr (((k) % BF_NGROUPS) == BF_GROUP && BF_DEV_KEEP(k)) {                    \
    if (cond) return fn.template run<__VA_ARGS__>();                                  \
  }

template <class F>
int bf_dispatch_model(const ModelConst& mc, F& fn) {
  const int id = mc.model_id, dx = mc.dx, dy = mc.dy;
  BF_ENTRY(0, id == 5 && dx == 8 && dy == 8, LgModel<8, 8>)
  BF_ENTRY(1, id == 2 && dy == 2, BotModel<2>)
  BF_ENTRY(2, id == 1, UngmModel)
  BF_ENTRY(3, id == 3, L63Model)
  BF_ENTRY(4, id == 5 && dx == 4 && dy == 4, LgModel<4, 4>)
  BF_ENTRY(5, id == 2 && dy == 1, BotModel<1>)
  BF_ENTRY(6, id == 5 && dx == 3 && dy == 3, LgModel<3, 3>)
  BF_ENTRY(7, id == 5 && dx == 4 && dy == 2, LgModel<4, 2>)
  BF_ENTRY(8, id == 6 && dx == 3, Sin1dModel<3>)
  BF_ENTRY(9, id == 5 && dx == 2 && dy == 2, LgModel<2, 2>)
  BF_ENTRY(10, id == 5 && dx == 3 && dy == 1, LgModel<3, 1>)
  BF_ENTRY(11, id == 6 && dx == 2, Sin1dModel<2>)
  BF_ENTRY(12, id == 5 && dx == 2 && dy == 1, LgModel<2, 1>)
  BF_ENTRY(13, id == 5 && dx == 1 && dy == 1, LgModel<1, 1>)
  BF_ENTRY(14, id == 6 && dx == 1, Sin1dModel<1>)
  BF_ENTRY(15, id == 7 && dx == 3, MsvModel<3>)
  BF_ENTRY(16, id == 7 && dx == 2, MsvModel<2>)
  BF_ENTRY(17, id == 7 && dx == 1, MsvModel<1>)
  (void)id; (void)dx; (void)dy;
  return -3;
}
