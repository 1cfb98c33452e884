// common.cuh -- shared device/host helpers for libbfilt (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define BF_MAXD 8          // max dx/dq/dy/dr of the register-resident ("small") models
#define BF_MAXTHETA 136    // LG dx=8,dy=8: 64 + 64 (+ spare)
#define BF_FULL 0xffffffffu

// Model constants, passed to kernels by value (__grid_constant__).  Everything a model's f/h/log-prob
// needs: the arrays of ParamsNLSSM/ParamsBPF (gaussfiltax/models.py:26-84) plus host-precomputed
// Cholesky factors (TFP's MultivariateNormalFullCovariance does cholesky(cov) at every call).
struct ModelConst {
  int model_id, dx, dq, dy, dr;
  float theta[BF_MAXTHETA];
  float q0[BF_MAXD], r0[BF_MAXD], m0[BF_MAXD];
  float Q[BF_MAXD * BF_MAXD], R[BF_MAXD * BF_MAXD], P0[BF_MAXD * BF_MAXD];
  float LQ[BF_MAXD * BF_MAXD];    // chol(Q), row-major dq x dq
  float LR[BF_MAXD * BF_MAXD];    // chol(R), row-major dr x dr
  float LP0[BF_MAXD * BF_MAXD];   // chol(P0)
  float iLR[BF_MAXD];             // 1 / diag(LR)
  float lp_const;                 // -sum(log diag LR) - 0.5*dy*log(2 pi)
};

// ---------------------------------------------------------------- Philox4x32-10 (counter-based RNG)
struct uint4x { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                      uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
  uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ __forceinline__ uint4x philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                      uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  uint4x r; r.x = c0; r.y = c1; r.z = c2; r.w = c3; return r;
}

// uniform in [0,1): top 24 bits, exactly representable, never 1.0
__host__ __device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
// uniform in (0,1]: for log()
__host__ __device__ __forceinline__ float u01_open0(uint32_t r) { return ((float)(r >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// Philox stream ids (counter word 3)
enum { BF_STREAM_INIT = 1, BF_STREAM_NOISE = 2, BF_STREAM_RESAMPLE = 3 };

#ifdef __CUDACC__
// two standard normals from two 32-bit words (Box-Muller).  Every operation is spelled out (one MUFU each for lg2, sqrt,
// sin, cos; explicit fma / mul) so that all kernels and all template instantiations draw bit-identical noise from the same
// counter, whatever the compiler would have contracted: a filter's result must not depend on which kernel variant ran it.
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float k = 1.0f / 16777216.0f;
  const float u1 = __fmaf_rn((float)(a >> 8), k, k);                 // (0,1]
  const float th = __fmul_rn((float)(b >> 8), 6.28318530717958647692f * k);   // [0, 2 pi)
  const float r = mufu_sqrt(__fmul_rn(mufu_lg2(u1), -1.38629436111989061883f));   // sqrt(-2 ln u1), ln = lg2 * ln 2
  z0 = __fmul_rn(r, mufu_cos(th)); z1 = __fmul_rn(r, mufu_sin(th));
}

// DQ standard normals for (trial b, particle i, step t): counter = (i_lo, i_hi | b<<8.., t, stream)
template <int DQ>
__device__ __forceinline__ void philox_normals(uint64_t seed, uint32_t stream, uint32_t t, uint32_t b, uint64_t i,
                                               float* z) {
#pragma unroll
  for (int k = 0; k < DQ; k += 4) {
    uint4x r = philox4x32((uint32_t)i, (uint32_t)(i >> 32) ^ (b << 4) ^ ((uint32_t)(k / 4) << 28), t, stream,
                          (uint32_t)seed, (uint32_t)(seed >> 32));
    float a0, a1, a2, a3;
    box_muller(r.x, r.y, a0, a1);
    z[k] = a0;
    if (k + 1 < DQ) z[k + 1] = a1;
    if (k + 2 < DQ) {
      box_muller(r.z, r.w, a2, a3);
      z[k + 2] = a2;
      if (k + 3 < DQ) z[k + 3] = a3;
    }
  }
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t stream, uint32_t t, uint32_t b, uint64_t i) {
  uint4x r = philox4x32((uint32_t)i, (uint32_t)(i >> 32) ^ (b << 4), t, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  return u01(r.x);
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(BF_FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BF_FULL, v, o);
  return v;
}
#endif  // __CUDACC__
