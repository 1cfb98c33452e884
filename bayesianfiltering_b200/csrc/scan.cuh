// scan.cuh -- prefix sum in the summation order of lax.associative_scan(add), the order jnp.cumsum
// has inside jax.random.choice (the reference's resampler, gaussfiltax/utils.py:210).
//
// Canonical order (oracle/scan.py): prefix[i] is the left-to-right sum, from the most- to the least-
// significant set bit of n = i+1, of the balanced pairwise-tree sums T(.) of the aligned power-of-two
// blocks that tile [0, n).  Equivalently, with E(node) the exclusive prefix of a tree node,
//     up-sweep    T(parent) = T(left) + T(right)
//     down-sweep  E(left) = E(parent),  E(right) = E(parent) + T(left)
//     prefix[i]   = E(leaf i+1)                       (so an EXCLUSIVE scan over n+1 leaves, leaf n = 0)
// fp32 addition is commutative but not associative, so the tree shape fixes every bit; an XOR butterfly
// realises the up-sweep in registers (both partners compute the same a+b), the values received at each
// butterfly level are exactly the T(left sibling) terms the down-sweep needs.  All adds are __fadd_rn so
// that ptxas cannot contract a neighbouring multiply into them.
//
// Tile = NWARPS*32 threads x 4 consecutive leaves per thread (levels 0-1 in the thread, 2-6 in the warp,
// 7.. across warps through shared memory).  Zero padding past the end is harmless: a padded block's sum
// is only ever added into prefixes of padded leaves.
#pragma once
#include "common.cuh"

template <int N>
struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <>
struct Log2<1> { static constexpr int v = 0; };

// Tree sum of the tile (up-sweep only).  Every thread returns the tile total.
// s_warp: NWARPS floats of shared memory.  Contains one __syncthreads.
template <int NWARPS>
__device__ __forceinline__ float tile_tree_sum(const float v[4], float* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float x = __fadd_rn(__fadd_rn(v[0], v[1]), __fadd_rn(v[2], v[3]));
#pragma unroll
  for (int j = 0; j < 5; ++j) x = __fadd_rn(x, __shfl_xor_sync(BF_FULL, x, 1 << j));
  if (lane == 0) s_warp[warp] = x;
  __syncthreads();
  float y = (lane < NWARPS) ? s_warp[lane] : 0.0f;
#pragma unroll
  for (int j = 0; j < Log2<NWARPS>::v; ++j) y = __fadd_rn(y, __shfl_xor_sync(BF_FULL, y, 1 << j));
  return __shfl_sync(BF_FULL, y, 0);
}

// Exclusive canonical scan of the tile.  e[k] = E(leaf 4*tid+k) given the tile's own exclusive prefix
// e_tile.  Returns the tile total.  Contains one __syncthreads (caller must make sure s_warp is free).
template <int NWARPS>
__device__ __forceinline__ float tile_exclusive_scan(const float v[4], float e_tile, float e[4], float* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float s01 = __fadd_rn(v[0], v[1]);
  const float s23 = __fadd_rn(v[2], v[3]);
  float x = __fadd_rn(s01, s23);
  float recv[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    recv[j] = __shfl_xor_sync(BF_FULL, x, 1 << j);
    x = __fadd_rn(x, recv[j]);
  }
  if (lane == 0) s_warp[warp] = x;
  __syncthreads();
  // every warp redoes the (tiny) cross-warp tree: avoids a second barrier
  constexpr int LW = Log2<NWARPS>::v;
  float y = (lane < NWARPS) ? s_warp[lane] : 0.0f;
  float rw[LW > 0 ? LW : 1];
#pragma unroll
  for (int j = 0; j < LW; ++j) {
    rw[j] = __shfl_xor_sync(BF_FULL, y, 1 << j);
    y = __fadd_rn(y, rw[j]);
  }
  const float total = __shfl_sync(BF_FULL, y, 0);
  float ew = e_tile;                       // lane l computes E(warp l)
#pragma unroll
  for (int j = LW - 1; j >= 0; --j)
    if ((lane >> j) & 1) ew = __fadd_rn(ew, rw[j]);
  float E = __shfl_sync(BF_FULL, ew, warp & 31);
#pragma unroll
  for (int j = 4; j >= 0; --j)
    if ((lane >> j) & 1) E = __fadd_rn(E, recv[j]);
  e[0] = E;
  e[1] = __fadd_rn(E, v[0]);
  const float e23 = __fadd_rn(E, s01);
  e[2] = e23;
  e[3] = __fadd_rn(e23, v[2]);
  return total;
}

// Canonical exclusive scan of n values by ONE CTA of NWARPS*32 threads, in place or out of place:
// out[i] = E(leaf i) for i in [0, n].  out must hold n+1 floats (out[n] = canonical total P(n)).
// Used for the grid-level pass over tile sums and by the resident (one CTA per trial) filter.
// s_chunk: shared scratch of at least ceil((n+1)/TILE) + NWARPS floats.
template <int NWARPS>
__device__ void cta_exclusive_scan(const float* in, float* out, int n, float* s_chunk) {
  constexpr int TILE = NWARPS * 128;
  const int tid = threadIdx.x;
  const int nleaf = n + 1;
  const int nchunk = (nleaf + TILE - 1) / TILE;
  float* s_warp = s_chunk + nchunk + 1;
  float v[4], e[4];
  if (nchunk == 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { int i = tid * 4 + k; v[k] = (i < n) ? in[i] : 0.0f; }
    __syncthreads();
    tile_exclusive_scan<NWARPS>(v, 0.0f, e, s_warp);
#pragma unroll
    for (int k = 0; k < 4; ++k) { int i = tid * 4 + k; if (i < nleaf) out[i] = e[k]; }
    __syncthreads();
    return;
  }
  // pass A: chunk tree sums
  for (int c = 0; c < nchunk; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { int i = c * TILE + tid * 4 + k; v[k] = (i < n) ? in[i] : 0.0f; }
    __syncthreads();                      // s_warp reuse
    float t = tile_tree_sum<NWARPS>(v, s_warp);
    if (tid == 0) s_chunk[c] = t;
  }
  __syncthreads();
  // pass B: canonical exclusive scan of the chunk sums (nchunk <= TILE, i.e. n < TILE^2)
#pragma unroll
  for (int k = 0; k < 4; ++k) { int i = tid * 4 + k; v[k] = (i < nchunk) ? s_chunk[i] : 0.0f; }
  __syncthreads();
  tile_exclusive_scan<NWARPS>(v, 0.0f, e, s_warp);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) { int i = tid * 4 + k; if (i < nchunk) s_chunk[i] = e[k]; }
  __syncthreads();
  // pass C: down-sweep inside each chunk
  for (int c = 0; c < nchunk; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { int i = c * TILE + tid * 4 + k; v[k] = (i < n) ? in[i] : 0.0f; }
    __syncthreads();
    tile_exclusive_scan<NWARPS>(v, s_chunk[c], e, s_warp);
#pragma unroll
    for (int k = 0; k < 4; ++k) { int i = c * TILE + tid * 4 + k; if (i < nleaf) out[i] = e[k]; }
  }
  __syncthreads();
}

// The reference's searchsorted (jnp.searchsorted side='left', method 'scan'): a FIXED bisection whose probe
// sequence is part of the contract because the tree-ordered fp32 cdf need not be monotone in its last bit
// (oracle/scan.py searchsorted_bisect).  `ex` is the EXCLUSIVE array (cdf[j] == ex[j+1]), n = len(cdf).
__device__ __forceinline__ int bisect_left_exclusive(const float* __restrict__ ex, int n, int n_levels, float r) {
  int low = 0, high = n;
  for (int l = 0; l < n_levels; ++l) {
    int mid = (low + high) >> 1;
    int m = mid < n - 1 ? mid : n - 1;
    bool go_left = r <= ex[m + 1];
    high = go_left ? mid : high;
    low = go_left ? low : mid;
  }
  return high < n - 1 ? high : n - 1;
}

__host__ __device__ __forceinline__ int bisect_levels(long long n) {
  int l = 0;                       // ceil(log2(n+1))
  while (((long long)1 << l) < n + 1) ++l;
  return l;
}
