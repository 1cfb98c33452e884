// pf_fused.cuh -- bootstrap particle filter, large-N path: ONE persistent cooperative kernel runs all T
// steps of gaussfiltax/inference.py:1330-1366 with the particle state streamed through HBM (sm_100a).
//
// Grid = one CTA of 32 warps per SM (cooperative launch; 16 warps at 128 registers for dx > 4).  A CTA owns a
// contiguous range of tiles (2048 particles, the unit of the canonical summation tree); its warps take the chunks
// (128 particles, 4 per lane) of that range round-robin, in the same order in every phase, so a run is
// bit-reproducible; no phase has a CTA barrier inside its loop.  A step is
//
//   A  ancestor of slot i from the PREVIOUS step's cdf (the resampling gather is fused into the propagate:
//      x = X[t-1][a_i], never materialised), x~ = f(x, q, u), ll = log p(y | x~), CTA partial of the online
//      (max, sum e, sum e^2, sum e x~) accumulator; the LAST CTA to finish folds all partials in a fixed order ->
//      max, Z, log-lik increment, ESS, resample flag, weighted mean (a sharded filter: every CTA of every rank folds,
//      same order, same bits)                                                                -- grid sync --
//   B  w = exp(ll - max) w_prev / Z ; no resampling: store w and go to the next step WITHOUT a grid sync;
//      resampling: w -> E, canonical tree sum of each chunk and tile of w                    -- grid sync --
//   C1 CTA 0: canonical exclusive scan of the tile sums                                      -- grid sync --
//   C2 one thread per tile: prefix and tau clamp interval of its 16 chunks; then the canonical exclusive scan
//      inside each chunk -> E[0..N] (cdf[j] = E[j+1], lax.associative_scan order: oracle/scan.py); for
//      power-of-two N the cdf is stored as its EFFECTIVE THRESHOLDS tau (see below)          -- grid sync --
//
// and after the last step one more gather materialises the carry.  HBM traffic of a resampling step with
// in-kernel Philox noise, per particle: A read x 4DX + write x~ 4DX + write ll 4 (+ tau window, cached);
// B read ll 4 + write w 4; C2 read w 4 + write tau 4  ->  8 DX + 20 bytes (52 B for DX = 4, as measured).
//
// The ancestor search is jnp.searchsorted's fixed bisection (oracle/scan.py searchsorted_bisect); its probe
// sequence is part of the contract because the tree-ordered fp32 cdf is not monotone in its last bit.  Running
// it literally costs 25 dependent loads per particle at N = 2^24 and was half of phase A.  Instead, for
// N = 2^k, phase C2 stores tau_j = the monotone thresholds with  bisect(cdf, r) == lower_bound(tau, r)  for every
// r (oracle/scan.py effective_thresholds_pow2: walking down the bisection tree, tau = min(v, .) under a left
// turn and max(v, .) under a right turn of the probe value v): the same 4 bytes per particle, after which ANY
// lower-bound search returns the reference's ancestors bit for bit.  Systematic positions are sorted: a chunk of
// 128 slots owns a window of tau (found by one bisection per chunk boundary at the start of the phase) and its
// answers are the running count of window entries below each slot's position (histogram + warp prefix sum, see
// ancestors4); multinomial positions look their cell up in a table of lower bounds rebuilt every step (begin_gather) or,
// for small N, bisect tau in two levels.  Other N keep the literal bisection on E.
//
// Compile-time experiment switches (never set in the product build, csrc/build.sh): BF_EXP_TRACE records a per-CTA timeline of
// one step (profiles/trace_summary.py); BF_EXP_NOSEARCH / NOLL / NOUNITRED and BF_EXP_MN_NOBUILD / NOLOOKUP / XONLY remove one
// piece of phase A to time the rest -- they compute WRONG filters by construction (profiles/README.md says which numbers came
// from them).
#pragma once
#include <cooperative_groups.h>
#include "pf_kernels.cuh"

namespace cg = cooperative_groups;

#define PF_FUSED_TILE_CHUNKS 16             // a tile = 16 chunks = 2048 particles: the unit whose sums one CTA scans
#ifndef PF_FUSED_WARPS
#define PF_FUSED_WARPS 32                   // warps per CTA, one CTA per SM (64 registers per thread)
#endif
#define PF_FUSED_WARPS_LARGE 16             // states with dx > 4: 16 warps at 128 registers per thread
#define PF_FUSED_WINCAP 256                 // floats of shared memory per warp for the histogram / tile tables / scan scratch
#define PF_FUSED_XCAP 160                   // particles a warp of a SHARDED filter stages when a chunk's ancestors live on a peer (two buffers)
#define PF_FUSED_MAXWB 32                    // window blocks (runs of consecutive owned tiles) per CTA, sharded filter with dealt blocks

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
  float* E;                                  // [ntiles*TILE] exclusive canonical cdf, E[N] = total; or tau[N] (use_tau)
  float* V;                                  // [ntiles] w[k*TILE]: cdf[k*TILE] = Etile[k] + V[k] are the probes above a tile
  float* total;                              // [1] canonical total (= E[N])
  int use_tau;                               // N = 2^k >= 8: E holds the effective thresholds tau
  float* tilesum;                            // [ntiles]
  float* Etile;                              // [ntiles+1]
  // Step statistics (max, sum e, sum e^2, sum e x) are reduced over a CANONICAL binary tree of the particle index, so that
  // their bits do not depend on the grid, the warp assignment or the number of GPUs a filter is sharded over:
  //   unit   512 consecutive particles (4 chunks, one warp, 16 particles per lane, accumulated in index order), reduced by
  //          the warp as soon as it is done; the partial goes to global memory with plain stores
  //   block  `bt` tiles = 4 bt units (bt = the largest power of two <= 8 dividing tiles_per_cta, so a CTA owns whole blocks):
  //          folded by a warp of the owning CTA after the CTA's end-of-phase barrier (balanced tree over the unit index)
  //   rank   the LAST CTA out of the phase (ticket) folds the rank's block roots (balanced tree) -> the step statistics on one
  //          GPU; in a sharded filter the rank's root goes to every peer and every CTA folds the `world` roots.
  PfPartial* upart;                          // [nunits] unit partials of this rank
  PfPartial* lroot;                          // [nroots] block roots of this rank
  PfPartial* broot;                          // [2][world] rank roots by step parity (sharded; in the exchange buffer)
  int nunits, bt, nroots;
  int off_root;                              // byte offset of broot inside an exchange buffer (sharded, p2p)
  StepStats* carry_stats;                    // [1] in (resume) / out
  const float* noise;                        // + (t*B + trial)*N*DQ, or NULL (Philox)
  const float* res_u;                        // multinomial + (t*B + trial)*N ; systematic + t*B + trial ; or NULL
  float* out_particles; float* out_weights; float* out_propagated; float* out_norm_weights; int* out_ancestors;
  float* out_mean; float* out_loglik; float* out_ess; int* out_resampled;      // already offset to (trial, t=0)
  float* carry_x; float* final_particles; float* final_weights;               // carry_w is W
  unsigned long long* phase_ns;             // [5] accumulated wall time of phases A, B, C1, C2, final as seen by CTA 0
  // ---- one filter sharded over `world` GPUs (power of two): this rank owns global particles [goff, goff + N).
  // Every index that leaves the shard (resampling positions, ancestors, tiles of the summation tree, Philox
  // counters) is GLOBAL; tau and the particle state of the peers are read through IPC-mapped pointers.
  int world, rank, goff, Ng, log2ng, log2ns; // world = 1: Ng = N, goff = 0
  unsigned int* fold_ctr;                    // ticket counter of the one-GPU statistics fold (zeroed by the host per run)
  StepStats* fold_stats;                     // its result
  float* tau_tile;                           // [N / 2048] last tau of each tile (one GPU, multinomial: two-level search)
  int* cell_lb;                              // [Ng + 1] one GPU, multinomial: lower bound of tau at the cell edges c * total / Ng, or NULL
  int coarse_log2;                           // log2 of its length, or -1: off
  int stage_ts;                              // C1 stages the tile sums in shared memory first (the host sized it for that)
  int c1_all, et_off;                        // every CTA scans the tile sums itself into shared memory at s_win + et_off (no barrier after C1)
  int tile_off, ntiles_g;                    // global index of local tile 0; tiles of the whole filter (tilesum/Etile/V are global)
  int phases;                                // PF_PH_* mask; PF_PH_ALL = the whole run in one launch (world = 1)
  // device tables [world] (in global memory: indexing a kernel parameter dynamically would spill it to local memory)
  const float* const* tau_peers;             // E of every rank (own entry = a.E)
  const float* const* xa_peers; const float* const* xb_peers;   // xbuf0 / xbuf1 of every rank
  // Exchange without NCCL (p2p = 1, phases = ALL): each rank owns an IPC-mapped exchange buffer holding
  // [flags 3 x 8 u64 | epoch | err | partials | tilesum | V]; a rank STORES its contribution straight into every peer's
  // buffer over NVLink and raises a sequence-numbered flag there; the readers spin on their own (local) flags.
  int p2p, off_part, off_ts, off_V;          // byte offsets of partials / tilesum / V inside an exchange buffer
  int xstage;                                // stage remote ancestor ranges in shared memory (needs the extra dynamic shared memory)
  // Bulk staging (sharded): before its units, a CTA copies the part of its ancestor range [s_lo[0], s_lo[last]] that lives on
  // other ranks -- tau and the particle state -- into these LOCAL buffers with every thread's loads in flight at once; the
  // units then never wait for an NVLink round trip (they cost the boundary CTAs 70 % when fetched chunk by chunk).
  float* stg_tau;                            // [gridDim.x][stg_blocks][stg_cap] staged thresholds of the ancestor brackets (sharded)
  int stg_cap, stg_blocks;
  int xpipe;                                 // sharded: prefetch the remote ancestor rows one chunk ahead (cp.async); off by default
  int strided, log2bt;                       // sharded: CTA c owns the blocks (bt tiles) c, c + grid, c + 2 grid, ... of its rank
  char* const* xchg_peers;                   // device table [world] of the exchange buffers (own entry = own buffer)
  unsigned long long* flags;                 // own flags [3][8]
  unsigned long long* epoch;                 // runs completed on this handle (device side: CUDA-graph replays need no new arguments)
  int* err;                                  // set when a peer did not show up within the spin budget
};

enum { PF_PH_A = 1, PF_PH_B = 2, PF_PH_C = 4, PF_PH_FINAL = 8, PF_PH_ALL = 15 };

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
// DQ % 4 == 0: the DQ normals of ONE particle (k-th of its quad) from DQ/4 calls -- nothing to share across the quad,
// and only one particle's noise is live at a time (large states would otherwise spill).
template <int DQ>
__device__ __forceinline__ void philox_particle_normals(unsigned long long seed, uint32_t stream, uint32_t t, uint32_t b,
                                                        uint32_t quad, uint32_t k, float* z) {
#pragma unroll
  for (int c = 0; c < DQ / 4; ++c) {
    const uint4x r = philox4x32(quad, b, t, stream | ((uint32_t)c << 8) | (k << 16), (uint32_t)seed, (uint32_t)(seed >> 32));
    box_muller(r.x, r.y, z[4 * c], z[4 * c + 1]);
    box_muller(r.z, r.w, z[4 * c + 2], z[4 * c + 3]);
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

// Unsorted (multinomial) positions on one GPU: two-level lower bound.  `s_last[k]` (shared memory) = the last tau of tile
// k, so the tile of the answer is found without touching global memory, and only the 11 levels inside that tile
// (2048 entries) are dependent global loads -- instead of log2(N) = 24.  tau is monotone, so any search order returns
// the reference's ancestor; the comparisons keep the bisection's form ("go right unless r <= v"), NaN included.
__device__ __forceinline__ void lower_bound4_two_level(const float* s_last, int log2nt, const float* tau, const float r[4],
                                                       int nm1, int anc[4]) {
  unsigned k0 = 0, k1 = 0, k2 = 0, k3 = 0;
  for (unsigned s = 1u << (log2nt - 1); s != 0; s >>= 1) {
    const float v0 = s_last[k0 + s], v1 = s_last[k1 + s], v2 = s_last[k2 + s], v3 = s_last[k3 + s];
    if (!(r[0] <= v0)) k0 += s;
    if (!(r[1] <= v1)) k1 += s;
    if (!(r[2] <= v2)) k2 += s;
    if (!(r[3] <= v3)) k3 += s;
  }
  k0 += (r[0] <= s_last[k0]) ? 0u : 1u; k1 += (r[1] <= s_last[k1]) ? 0u : 1u;
  k2 += (r[2] <= s_last[k2]) ? 0u : 1u; k3 += (r[3] <= s_last[k3]) ? 0u : 1u;
  const unsigned nt = 1u << log2nt;
  constexpr unsigned TILE = PF_FUSED_TILE_CHUNKS * 128;
  // a position above every tile's last threshold: the in-tile search runs on the last tile and the clamp gives N - 1
  const float* t0 = tau + (size_t)min(k0, nt - 1) * TILE; const float* t1 = tau + (size_t)min(k1, nt - 1) * TILE;
  const float* t2 = tau + (size_t)min(k2, nt - 1) * TILE; const float* t3 = tau + (size_t)min(k3, nt - 1) * TILE;
  unsigned l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  for (unsigned s = TILE / 2; s != 0; s >>= 1) {
    const float v0 = t0[l0 + s], v1 = t1[l1 + s], v2 = t2[l2 + s], v3 = t3[l3 + s];
    if (!(r[0] <= v0)) l0 += s;
    if (!(r[1] <= v1)) l1 += s;
    if (!(r[2] <= v2)) l2 += s;
    if (!(r[3] <= v3)) l3 += s;
  }
  l0 += (r[0] <= t0[l0]) ? 0u : 1u; l1 += (r[1] <= t1[l1]) ? 0u : 1u;
  l2 += (r[2] <= t2[l2]) ? 0u : 1u; l3 += (r[3] <= t3[l3]) ? 0u : 1u;
  anc[0] = min((int)(min(k0, nt - 1) * TILE + l0), nm1); anc[1] = min((int)(min(k1, nt - 1) * TILE + l1), nm1);
  anc[2] = min((int)(min(k2, nt - 1) * TILE + l2), nm1); anc[3] = min((int)(min(k3, nt - 1) * TILE + l3), nm1);
}

// ONE query, two levels: the tile from the table of last thresholds in shared memory, then the 11 levels inside it.
// Same result as lower_bound_tau_pow2 on the monotone tau (NaN r: n), 11 dependent global loads instead of log2(n).
__device__ __forceinline__ int lower_bound_two_level(const float* s_last, int log2nt, const float* tau, float r) {
  unsigned k = 0;
  for (unsigned s = 1u << (log2nt - 1); s != 0; s >>= 1)
    if (!(r <= s_last[k + s])) k += s;
  k += (r <= s_last[k]) ? 0u : 1u;
  constexpr unsigned TILE = PF_FUSED_TILE_CHUNKS * 128;
  if (k >= (1u << log2nt)) return (int)(TILE << log2nt);
  const float* t = tau + (size_t)k * TILE;
  unsigned l = 0;
  for (unsigned s = TILE / 2; s != 0; s >>= 1)
    if (!(r <= t[l + s])) l += s;
  l += (r <= t[l]) ? 0u : 1u;
  return (int)(k * TILE + l);
}

// fixed bisection / lower bound of ONE query on the monotone thresholds tau[0..2^log2n); result in [0, n]
__device__ __forceinline__ int lower_bound_tau_pow2(const float* tau, int log2n, float r) {
  unsigned low = 0;
  for (unsigned s = 1u << (log2n - 1); s != 0; s >>= 1)
    if (!(r <= tau[low + s])) low += s;
  return (int)(low + ((r <= tau[low]) ? 0u : 1u));
}

// number of entries of win[p0..wn) that are < r (NaN r counts every entry, like the bisection's "go right")
__device__ __forceinline__ int lower_bound_window(const float* win, int p0, int wn, float r) {
  int pos = p0, len = wn - p0;
  while (len > 0) {
    const int half = len >> 1;
    if (!(r <= win[pos + half])) { pos += half + 1; len -= half + 1; } else len = half;
  }
  return pos;
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// Canonical tree sum of one warp-chunk (128 consecutive leaves, 4 per lane); every lane returns it.
__device__ __forceinline__ float chunk_tree_sum(const float v[4]) {
  float x = __fadd_rn(__fadd_rn(v[0], v[1]), __fadd_rn(v[2], v[3]));
#pragma unroll
  for (int j = 0; j < 5; ++j) x = __fadd_rn(x, __shfl_xor_sync(BF_FULL, x, 1 << j));
  return x;
}

// Exclusive canonical scan of one warp-chunk inside its tile, barrier-free: the tree sums of the tile's NWARPS
// chunks (s_wsum) were left in shared memory by phase B.  Same additions in the same order as
// tile_exclusive_scan (scan.cuh).  wpos = position of this chunk in the tile.
// With TAU it also turns the chunk's cdf into its effective thresholds: the bisection tree over the boundaries
// j of the tile is walked from the tile's node down; the probe value of a node with middle m is cdf[m] = E[m+1]:
//   across chunks (half-width >= 128 leaves): E(chunk start) + v[first leaf of that chunk] (s_v0) -> lane l of every warp
//   inside a chunk (half-width 64 .. 4)     : e[1] of the lane that owns leaf m                    -> one shuffle per level
//   inside a thread (half-width 2, 1)       : e[3], e[2] and the next thread's e[0].
// (lo, hi) is the clamp interval the levels ABOVE the tile impose; ncap = min(TILE, N) (a power of two).
// In-chunk part: `x`/`recv` are the warp's up-sweep (chunk_upsweep), E_chunk the canonical prefix at the chunk's first
// leaf, nxtw the prefix at the NEXT chunk's first leaf, (lo, hi) the clamp interval every level above the chunk imposes.
__device__ __forceinline__ void chunk_upsweep(const float v[4], float& s01, float recv[5]) {
  s01 = __fadd_rn(v[0], v[1]);
  float x = __fadd_rn(s01, __fadd_rn(v[2], v[3]));
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    recv[j] = __shfl_xor_sync(BF_FULL, x, 1 << j);
    x = __fadd_rn(x, recv[j]);
  }
}
template <bool TAU>
__device__ __forceinline__ void chunk_downsweep(const float v[4], float s01, const float recv[5], float E_chunk, float nxtw,
                                                float lo, float hi, int ncap, float e[4], float tau[4]) {
  const int lane = threadIdx.x & 31;
  float E = E_chunk;
#pragma unroll
  for (int j = 4; j >= 0; --j)
    if ((lane >> j) & 1) E = __fadd_rn(E, recv[j]);
  e[0] = E;
  e[1] = __fadd_rn(E, v[0]);
  const float e23 = __fadd_rn(E, s01);
  e[2] = e23;
  e[3] = __fadd_rn(e23, v[2]);
  if (!TAU) return;
#pragma unroll
  for (int sl = 16; sl >= 1; sl >>= 1) {                       // nodes spanning 2*sl lanes
    const int midl = (lane & ~(2 * sl - 1)) + sl;
    const float c = clampf(__shfl_sync(BF_FULL, e[1], midl), lo, hi);
    if (2 * sl * 4 <= ncap) { if (lane < midl) hi = c; else lo = c; }
  }
  float nxt = __shfl_down_sync(BF_FULL, e[0], 1);              // E of the next thread's first leaf
  if (lane == 31) nxt = nxtw;
  const float c2 = clampf(e[3], lo, hi);                       // node (j0, j0+4), middle j0+2
  const float c1 = clampf(e[2], lo, c2);                       // node (j0, j0+2), middle j0+1
  const float c3 = clampf(nxt, c2, hi);                        // node (j0+2, j0+4), middle j0+3
  tau[0] = clampf(e[1], lo, c1);
  tau[1] = clampf(e[2], c1, c2);
  tau[2] = clampf(e[3], c2, c3);
  tau[3] = clampf(nxt, c3, hi);
}

template <int NWARPS, bool TAU>
__device__ __forceinline__ void chunk_scan(const float v[4], int wpos, const float* s_wsum, const float* s_v0, float e_tile,
                                           float e_next_tile, float lo, float hi, int ncap, float e[4], float tau[4]) {
  const int lane = threadIdx.x & 31;
  float s01, recv[5];
  chunk_upsweep(v, s01, recv);
  constexpr int LW = Log2<NWARPS>::v;
  float y = (lane < NWARPS) ? s_wsum[lane] : 0.0f;
  float rw[LW > 0 ? LW : 1];
#pragma unroll
  for (int j = 0; j < LW; ++j) {
    rw[j] = __shfl_xor_sync(BF_FULL, y, 1 << j);
    y = __fadd_rn(y, rw[j]);
  }
  float ew = e_tile;                       // lane l: E(first leaf of chunk l)
#pragma unroll
  for (int j = LW - 1; j >= 0; --j)
    if ((lane >> j) & 1) ew = __fadd_rn(ew, rw[j]);
  const float E_chunk = __shfl_sync(BF_FULL, ew, wpos);
  float nxtw = 0.0f;
  if (TAU) {
    const float cw = __fadd_rn(ew, (lane < NWARPS) ? s_v0[lane] : 0.0f);     // lane l: cdf[first boundary of chunk l]
#pragma unroll
    for (int swd = NWARPS / 2; swd >= 1; swd >>= 1) {            // nodes spanning 2*swd chunks
      const int midw = (wpos & ~(2 * swd - 1)) + swd;
      const float c = clampf(__shfl_sync(BF_FULL, cw, midw & 31), lo, hi);
      if (2 * swd * 128 <= ncap) { if (wpos < midw) hi = c; else lo = c; }
    }
    nxtw = (wpos + 1 < NWARPS) ? __shfl_sync(BF_FULL, ew, (wpos + 1) & 31) : e_next_tile;
  }
  chunk_downsweep<TAU>(v, s01, recv, E_chunk, nxtw, lo, hi, ncap, e, tau);
}

// The tile-level part of chunk_scan<NWARPS, true> for all NWARPS chunks of one tile, by ONE thread (phase C2 does it once
// per tile instead of once per lane and chunk): the same additions and clamps in the same order.
//   cE[w] = prefix at the first leaf of chunk w (cE[NWARPS] = e_next_tile), clo/chi[w] = clamp interval of chunk w.
template <int NWARPS>
__device__ __forceinline__ void tile_levels(const float* s_wsum, const float* s_v0, float e_tile, float e_next_tile, float lo,
                                            float hi, int ncap, float* cE, float* clo, float* chi) {
  constexpr int LW = Log2<NWARPS>::v;
  float S[LW + 1][NWARPS];                  // S[j][b]: tree sum of block b of 2^j chunks
#pragma unroll
  for (int l = 0; l < NWARPS; ++l) S[0][l] = s_wsum[l];
#pragma unroll
  for (int j = 1; j <= LW; ++j)
#pragma unroll
    for (int b = 0; b < (NWARPS >> j); ++b) S[j][b] = __fadd_rn(S[j - 1][2 * b], S[j - 1][2 * b + 1]);
  float ew[NWARPS], cw[NWARPS];
#pragma unroll
  for (int l = 0; l < NWARPS; ++l) {
    float x = e_tile;
#pragma unroll
    for (int j = LW - 1; j >= 0; --j)
      if ((l >> j) & 1) x = __fadd_rn(x, S[j][(l >> j) ^ 1]);
    ew[l] = x;
    cw[l] = __fadd_rn(x, s_v0[l]);
  }
#pragma unroll
  for (int w = 0; w < NWARPS; ++w) {
    float l = lo, h = hi;
#pragma unroll
    for (int swd = NWARPS / 2; swd >= 1; swd >>= 1) {
      const int midw = (w & ~(2 * swd - 1)) + swd;
      const float c = clampf(cw[midw & (NWARPS - 1)], l, h);
      if (2 * swd * 128 <= ncap) { if (w < midw) h = c; else l = c; }
    }
    cE[w] = ew[w]; clo[w] = l; chi[w] = h;
  }
  cE[NWARPS] = e_next_tile;
}

// Canonical exclusive scan of n (+1 leaf for the total) values already staged in shared memory, n + 1 <= 16384:
// warps take the 128-leaf blocks in turn (tree sums), warp 0 scans the <= 128 block sums as the leaves of the upper
// tree, then the blocks are swept down from their prefixes.  Two barriers instead of cta_exclusive_scan's dozens; the
// same zero-padded binary tree, so the same bits.  s_blk: 256 floats of scratch.  out[0 .. n] (global).
template <int CW>
__device__ __forceinline__ void staged_tile_scan(const float* s_in, float* out, int n, float* s_blk) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblk = (n + 1 + 127) >> 7;                 // <= 128
  for (int b = warp; b < 128; b += CW) {
    float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (b < nblk) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int i = b * 128 + lane * 4 + k; v[k] = (i < n) ? s_in[i] : 0.0f; }
    }
    const float t = chunk_tree_sum(v);
    if (lane == 0) s_blk[b] = t;
  }
  __syncthreads();
  if (warp == 0) {
    const float v[4] = {s_blk[lane * 4], s_blk[lane * 4 + 1], s_blk[lane * 4 + 2], s_blk[lane * 4 + 3]};
    float s01, recv[5], e[4], unused[4];
    chunk_upsweep(v, s01, recv);
    chunk_downsweep<false>(v, s01, recv, 0.0f, 0.0f, 0.0f, 0.0f, 0, e, unused);
#pragma unroll
    for (int k = 0; k < 4; ++k) s_blk[128 + lane * 4 + k] = e[k];
  }
  __syncthreads();
  for (int b = warp; b < nblk; b += CW) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = b * 128 + lane * 4 + k; v[k] = (i < n) ? s_in[i] : 0.0f; }
    float s01, recv[5], e[4], unused[4];
    chunk_upsweep(v, s01, recv);
    chunk_downsweep<false>(v, s01, recv, s_blk[128 + b], 0.0f, 0.0f, 0.0f, 0, e, unused);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = b * 128 + lane * 4 + k; if (i <= n) out[i] = e[k]; }
  }
}

// The particle state is read once and written once per step and is far larger than L2: stream it (evict-first), so
// that the per-particle scalars that ARE re-read within a step (log-likelihoods, weights, tau: 4 B each) stay in L2.
template <int DX>
__device__ __forceinline__ void store_state_cs(float* base, size_t i, const float* x) {
  if constexpr (DX % 4 == 0) {
    float4* p = reinterpret_cast<float4*>(base) + i * (DX / 4);
#pragma unroll
    for (int k = 0; k < DX / 4; ++k) __stcs(p + k, make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]));
  } else if constexpr (DX % 2 == 0) {
    float2* p = reinterpret_cast<float2*>(base) + i * (DX / 2);
#pragma unroll
    for (int k = 0; k < DX / 2; ++k) __stcs(p + k, make_float2(x[2 * k], x[2 * k + 1]));
  } else {
#pragma unroll
    for (int k = 0; k < DX; ++k) __stcs(base + i * DX + k, x[k]);
  }
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

// Cold piece of the canonical statistics tree, kept out of line so that its stack does not weigh on the hot loop.
// balanced tree over the aligned range [first, first + R) of block roots (R a power of two > 2; N >= 2^28 only), by binary carries
template <int DX>
__device__ __noinline__ void fold_root_range(const PfPartial* roots, int first, int R, int nroots, PfPartial* out) {
  OnlineAcc<DX> st[12], cur;
  for (int i = 0; i < R; ++i) {
    cur.init();
    if (first + i < nroots) cur.load(roots + first + i);
    int lv = 0;
    for (int bits = i; bits & 1; bits >>= 1, ++lv) { OnlineAcc<DX> l = st[lv]; l.merge_acc(cur); cur = l; }
    st[lv] = cur;
  }
  cur.store(out);
}

// SH: the sharded-filter code paths (peer tables, staging of remote ancestor ranges, cross-GPU barriers) are compiled
// only into the SH = true instantiation; the one-GPU kernel does not carry their registers.
template <class M, int CW, int CTAS_PER_SM, bool SH, bool LEAN>
__global__ void __launch_bounds__(CW * 32, CTAS_PER_SM)
k_pf_fused(const __grid_constant__ ModelConst mc, const __grid_constant__ PfFusedArgs a_in) {
  // LEAN: the throughput configuration -- in-kernel Philox noise and uniforms, systematic resampling on the tau path,
  // whole run in one launch, summaries only (mean / log-lik / ESS / flag per step, final state).  Its optional
  // pointers are compile-time NULL here, which removes their code from the hot loop (measured: phase A 298 -> 272 us
  // at N = 2^24; the kernel is instruction-issue- and instruction-fetch-bound).  The host picks it when the call
  // qualifies (bfilt_api.cu), otherwise the general instantiation runs.
  PfFusedArgs a = a_in;
  if constexpr (LEAN) {
    a.noise = nullptr; a.res_u = nullptr; a.out_particles = nullptr; a.out_propagated = nullptr; a.out_ancestors = nullptr;
    a.out_norm_weights = nullptr; a.out_weights = nullptr; a.resampler = 1; a.use_tau = 1; a.phases = PF_PH_ALL;
  }
  constexpr int DX = M::DX, DQ = M::DQ, DY = M::DY;
  constexpr int TW = PF_FUSED_TILE_CHUNKS;              // chunks per tile: a level of the canonical tree, not a CTA shape
  constexpr int NT = CW * 32, TILE = TW * 128, WINCAP = PF_FUSED_WINCAP;
  cg::grid_group grid = cg::this_grid();
  // dynamic shared memory (words): [TW*tpc] chunk tree sums | [TW*tpc] first weight of each chunk |
  // [2*tpc] clamp interval of each tile | [TW*tpc+2] tau window starts | [CW*WINCAP] tau windows (phase A),
  // aliased by CTA 0's scan scratch (C1)
  extern __shared__ __align__(16) float s_dyn[];
  const int tpc = a.tiles_per_cta;
  float* s_wsum = s_dyn;
  float* s_v0 = s_wsum + TW * tpc;
  float* s_meta_f = s_v0 + TW * tpc;
  int* s_lo = reinterpret_cast<int*>(s_meta_f + 2 * tpc);       // [(TW+1)*tpc + 2] tau window start of each chunk (+ one end per window block)
  float* s_win = s_dyn + (((3 * TW + 3) * tpc + 2 + 3) & ~3);   // 16-byte aligned (vector accesses to the histogram)
  float* s_xst = s_win + CW * WINCAP;                           // [CW][2][XCAP*DX], sharded filters only (host sizes it)
  float* s_et = s_win + a.et_off;                               // [ntiles_g + 1] tile prefixes of the step (c1_all; aliases s_xst, idle in phase C)
  const bool c1_all = a.c1_all != 0;
  auto etile = [&](int i) -> float { return c1_all ? s_et[i] : __ldcg(&a.Etile[i]); };
  // the particles of a remote ancestor window are fetched ONE CHUNK AHEAD with cp.async (16-byte rows): the NVLink round
  // trip (3-10 us under load) overlaps the propagation of the chunk in hand instead of stalling the warp
  constexpr bool XPIPE = SH && (DX == 4);
  bool x_pre = false;                                 // the chunk's window is (being) staged by the pipeline
  int x_buf = 0;                                      // ... in this buffer of the warp
  __shared__ __align__(16) StepStats s_stats;
  __shared__ int s_last;                              // this CTA took the last ticket of the step (it folds the statistics)
  __shared__ int s_ctr[4];                            // chunk queues of phases A / B / C2
  __shared__ int s_br[2 * PF_FUSED_MAXWB];            // sharded filter: bracket [lo, hi) of the ancestors of each window block (begin_gather)
  __shared__ int s_stg[PF_FUSED_MAXWB];               // ... its remote thresholds are staged in this CTA's part of a.stg_tau

  const int N = a.N, T = a.T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  // Ownership.  One GPU: CTA c owns the tiles [c tpc, (c + 1) tpc).  Sharded filter (a.strided): CTA c owns the BLOCKS (bt
  // consecutive tiles, the unit of the statistics tree) c, c + G, c + 2G, ... of its rank -- the slots whose ancestors live
  // on another rank are a contiguous part of the rank's range, and dealing the blocks round gives every CTA the same share
  // of them (contiguous ownership left a few CTAs with nothing but NVLink round trips, and everyone else waiting).
  // Nothing that is computed depends on who owns what: every reduction is a fixed tree over the particle index.
  const bool strided = SH && a.strided != 0;
  const int log2bt = a.log2bt, btm = (1 << log2bt) - 1;
  const int tile0 = blockIdx.x * tpc;                  // (contiguous ownership)
  int ntile_cta = max(0, min(tile0 + tpc, a.ntiles) - tile0);
  if (strided) {
    const int nblk_rank = a.ntiles >> log2bt;          // (the host makes ntiles a multiple of bt)
    ntile_cta = ((int)blockIdx.x < nblk_rank) ? (((nblk_rank - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) << log2bt) : 0;
  }
  auto tile_of = [&](int tl) -> int {                  // tile tl of this CTA -> tile of the rank
    return strided ? ((((tl >> log2bt) * (int)gridDim.x + (int)blockIdx.x) << log2bt) + (tl & btm)) : tile0 + tl;
  };
  auto pbase = [&](int q) -> int { return tile_of(q / TW) * TILE + (q % TW) * 128; };     // first particle of chunk q of this CTA
  // chunks (128 consecutive particles) of this CTA; warps take them
  // from a shared queue (the SM's warp arbiter favours some warps: a static split leaves the others as a tail)
  const int nchunk_cta = TW * ntile_cta;
  const int cta_base = tile0 * TILE;                   // (contiguous ownership)
  const int nchunk_valid = strided ? nchunk_cta : max(0, min(nchunk_cta, (N - cta_base + 127) / 128));
  // window blocks: runs of consecutive chunks, whose tau windows share their end points -- the whole CTA, or one block
  const int wb = strided ? (TW << log2bt) : max(nchunk_cta, 1);
  const int nwb = strided ? (ntile_cta >> log2bt) : 1;
  auto sidx = [&](int q) -> int { return strided ? (q / wb) * (wb + 1) + (q % wb) : q; };
#ifdef BF_DYNAMIC_CHUNKS
  auto next_chunk = [&](int which) -> int {
    int q = 0;
    if (lane == 0) q = atomicAdd(&s_ctr[which], 1);
    return __shfl_sync(BF_FULL, q, 0);
  };
  auto next_chunks4 = [&](int which) -> int {          // four consecutive chunks: their loads are issued together
    int q = 0;
    if (lane == 0) q = atomicAdd(&s_ctr[which], 4);
    return __shfl_sync(BF_FULL, q, 0);
  };
#else
  // Static round-robin: warp w takes chunks w, w + CW, ... -- the order in which a thread accumulates its
  // statistics is then fixed, so a run is bit-reproducible (and a run split in time equals the unsplit run).
  int turn[3] = {0, 0, 0};
  auto next_chunk = [&](int which) -> int { return warp + CW * turn[which]++; };
  auto next_chunks4 = [&](int which) -> int { return 4 * (warp + CW * turn[which]++); };
#endif
  if (tid < 4) s_ctr[tid] = 0;
  __syncthreads();
  const int Ng = a.Ng;                                 // particles of the whole filter (= N unless sharded)
  const float invN = 1.0f / (float)Ng;
  const int levels = bisect_levels(N);
  const bool use_tau = a.use_tau != 0;
  const bool all_phases = a.phases == PF_PH_ALL;
  const int ns_mask = (1 << (a.log2ns < 0 ? 0 : a.log2ns)) - 1;
  // global accessors: entry j of the (virtual) global tau array, particle j of the previous step's state
  bool stg_on = false;                                // the remote thresholds of the window block in hand are staged locally (sharded)
  int stg_lo = 0, stg_hi = -1;
  const float* stg_cta = (SH && a.stg_tau) ? a.stg_tau + (size_t)blockIdx.x * a.stg_blocks * a.stg_cap : nullptr;
  const float* stg_tau_c = stg_cta;
  auto set_stage = [&](int b) {                        // window block b of this CTA: where tauG finds its remote thresholds
    if constexpr (SH) { stg_on = s_stg[b] != 0; stg_lo = s_br[2 * b]; stg_hi = s_br[2 * b + 1] - 1; stg_tau_c = stg_cta + (size_t)b * a.stg_cap; }
  };
  auto tauG = [&](int j) -> float {
    if constexpr (!SH) return a.E[j];
    else {
      if ((j >> a.log2ns) == a.rank) return a.E[j & ns_mask];
      if (stg_on && j >= stg_lo && j <= stg_hi) return __ldcg(stg_tau_c + (j - stg_lo));
      return a.tau_peers[j >> a.log2ns][j & ns_mask];
    }
  };
  const bool w_vec = (reinterpret_cast<size_t>(a.W) & 15) == 0;     // carry of trial b starts at b*N floats
  const bool use_win = use_tau && a.resampler == 1;      // sorted positions: windowed search
  const bool use_cells = !SH && use_tau && a.resampler != 1 && a.cell_lb != nullptr && a.log2ng >= 0;   // unsorted positions: cell table
  // (i + u) / 2^k is an exact scaling: multiplying by 2^-k gives the same bits as the oracle's division
  const float invn_pow2 = (a.log2ng >= 0) ? __int_as_float((127 - a.log2ng) << 23) : 0.0f;
  // ---- cross-GPU barrier of the persistent sharded kernel.  Every thread that stored into a peer's buffer has run
  // __threadfence_system() and the grid has synchronised before CTA 0 raises the flags; readers spin on local memory.
  const bool p2p = SH && a.p2p != 0 && a.world > 1;
  const unsigned long long seq0 = p2p ? (*a.epoch) * (3ull << 22) : 0ull;
  auto xbarrier = [&](int ph, int step) {
    const unsigned long long seq = seq0 + 3ull * (unsigned long long)step + (unsigned long long)ph + 1ull;
    if (blockIdx.x == 0 && tid < a.world && tid != a.rank) {
      __threadfence_system();
      volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(a.xchg_peers[tid]) + ph * 8 + a.rank;
      *f = seq;
    }
    if (tid < a.world && tid != a.rank) {
      volatile unsigned long long* f = a.flags + ph * 8 + tid;
      const unsigned long long t0 = globaltimer_ns();
      while (*f < seq) {
        if (globaltimer_ns() - t0 > 4000000000ull) { *a.err = 1; break; }     // 4 s: a peer died; do not hang the GPU
      }
      __threadfence_system();          // acquire; a fence of scope >= cluster also drops this SM's stale L1 lines
    }
    __syncthreads();
  };

  // the carry entering the call: materialised particles (gather = false) for a whole run; for a single phase of a
  // sharded step the host keeps carry_stats current between the launches
  bool w_uni = (a.resume || !all_phases) ? (a.carry_stats->w_uniform != 0) : (a.w_uniform0 != 0);
  bool gather = all_phases ? false : (a.carry_stats->flag != 0);
  const float* x_prev = a.x_init;
  float res_total = 0.0f, res_u0 = 0.0f, res_scale = 0.0f;   // of the resampling being gathered (scale = Ng / total)
  const float* const* xp_stage = nullptr;             // peers' copies of the state being gathered (sharded filter)

  // systematic position of a slot when Ng is a power of two (the tau path): exact scaling, then one rounding
  auto pos_sys = [&](int slot, float u) -> float { return __fmul_rn(res_total, __fmul_rn(__fadd_rn((float)slot, u), invn_pow2)); };
  auto position = [&](int slot, float u) -> float {
    if (a.log2ng >= 0)
      return (a.resampler == 1) ? __fmul_rn(res_total, __fmul_rn(__fadd_rn((float)slot, u), invn_pow2))
                                : __fmul_rn(res_total, __fsub_rn(1.0f, u));
    return resample_position(a.resampler, res_total, u, slot, Ng);
  };

  // Lower bounds of the 128 SORTED positions (s0 + i + u0) * total / Ng, i < 128, by counting: out[k] belongs to slot
  // s0 + 4 lane + k; [lo, lo + wn] brackets the answers.  Warp-synchronous; wn <= 32 * 128.
  auto count_window = [&](int lo, int wn, int s0, float u0, int out[4]) {
    // Counting instead of searching.  Slot i's answer is lo + #{j in the window : tau_j < r_i}; positions are
    // non-decreasing in i, so tau_j is below the positions of exactly the slots i >= k_j, k_j = first slot of the
    // chunk whose position exceeds tau_j.  Every lane turns its share of the window (about one tau per slot) into
    // k_j -- an estimate from the inverse of the position formula, corrected against the formula itself, so the
    // result is exact whatever the rounding of the estimate -- and counts it in a 129-bin histogram in shared
    // memory; a warp prefix sum of the histogram gives all 128 answers.  ~30 instructions per slot where the
    // bounded bisections took ~75, and no staging of the window.
    int* hist = reinterpret_cast<int*>(s_win) + warp * 160;
    *reinterpret_cast<int4*>(hist + lane * 4) = make_int4(0, 0, 0, 0);
    if (lane == 0) hist[128] = 0;
    __syncwarp();
    const float fs0 = (float)s0;                    // a multiple of 128 below 2^31: exact
    for (int j0 = 0; j0 < wn; j0 += 128) {          // four independent loads in flight per lane
      float tv[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) { const int i = j0 + lane + 32 * m; tv[m] = (i < wn) ? tauG(lo + i) : 0.0f; }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        if (j0 + lane + 32 * m < wn) {
          const float tj = tv[m];
          // estimate: pos_sys(s, u) > tj  <=>  s > tj * Ng / total - u
          float tr = __fsub_rn(__fsub_rn(__fmul_rn(tj, res_scale), u0), fs0);
          tr = fminf(fmaxf(tr, -1.0f), 128.0f);     // (NaN -> -1)
          int g = min((int)floorf(tr) + 1, 128);
          while (g > 0 && pos_sys(s0 + g - 1, u0) > tj) --g;
          while (g < 128 && !(pos_sys(s0 + g, u0) > tj)) ++g;
          atomicAdd(hist + g, 1);
        }
      }
    }
    __syncwarp();
    const int4 h4 = *reinterpret_cast<const int4*>(hist + lane * 4);
    const int c0 = h4.x, c1 = c0 + h4.y, c2 = c1 + h4.z, c3 = c2 + h4.w;
    int incl = c3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(BF_FULL, incl, o);
      if (lane >= o) incl += up;
    }
    const int ex = lo + incl - c3;
    out[0] = min(ex + c0, Ng - 1); out[1] = min(ex + c1, Ng - 1); out[2] = min(ex + c2, Ng - 1); out[3] = min(ex + c3, Ng - 1);
    if (!(res_total == res_total)) {               // NaN total: every comparison of the bisection fails -> last index
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k] = Ng - 1;
    }
    __syncwarp();                                    // the histogram is reused by this warp's next chunk
  };

  // Once per phase that gathers: total and the systematic uniform.  ts = local step whose resampling is gathered.
  auto begin_gather = [&](int ts) {
    stg_on = false;                                  // (the staged copies of the previous step are stale)
    res_total = __ldcg(a.total);
    res_scale = __fdiv_rn((float)Ng, res_total);
    if (a.resampler == 1)
      res_u0 = a.res_u ? a.res_u[(ptrdiff_t)ts * a.B]
                       : philox_uniform(a.seed, BF_STREAM_RESAMPLE, (uint32_t)(ts + a.t_offset), (uint32_t)a.trial, 0ull);
    if (!SH && use_cells) {
      // Unsorted (multinomial) positions, one GPU: first the CELL TABLE  cell_lb[c] = lower_bound(tau, c * total / Ng),
      // c = 0 .. Ng -- the systematic search with u = 0, i.e. sorted queries answered by the counting scheme at streaming
      // cost.  A position r = total * w, w = 1 - u in [0, 1], lies between the edges of cell c = floor(w * Ng) (computed
      // exactly: Ng is a power of two) and both are rounded by the same monotone multiplication, so its lower bound lies in
      // [cell_lb[c], cell_lb[c + 1]]: two adjacent table entries and a search over the ~1 tau of the cell replace the 11
      // dependent probes of a tile (and when w * Ng is an integer, r IS the edge and cell_lb[c] is the answer).
      const bool two = a.coarse_log2 >= 0;
      if (two) {
        for (int k = tid; k < (1 << a.coarse_log2); k += NT) s_win[k] = __ldcg(a.tau_tile + k);
        __syncthreads();
      }
      for (int q = tid; q <= nchunk_valid; q += NT) {
        const int slot = (q < nchunk_valid) ? cta_base + 128 * q : min(cta_base + 128 * nchunk_cta, N);
        const float r = pos_sys(slot, 0.0f);
        s_lo[q] = (nchunk_valid <= 0) ? 0 : two ? lower_bound_two_level(s_win, a.coarse_log2, a.E, r) : lower_bound_tau_pow2(a.E, a.log2ng, r);
      }
      __syncthreads();
#ifndef BF_EXP_MN_NOBUILD
      for (int q = warp; q < nchunk_valid; q += CW) {
        const int lo = s_lo[q], wn = s_lo[q + 1] - lo, s0 = cta_base + 128 * q;
        int lb[4];
        if (wn <= 32 * 128) count_window(lo, wn, s0, 0.0f, lb);
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) lb[k] = min(lower_bound_tau_pow2(a.E, a.log2ng, pos_sys(s0 + 4 * lane + k, 0.0f)), Ng - 1);
        }
        const int c = s0 + 4 * lane;
        if (c + 3 < N) *reinterpret_cast<int4*>(a.cell_lb + c) = make_int4(lb[0], lb[1], lb[2], lb[3]);
        else {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (c + k < N) a.cell_lb[c + k] = lb[k];
        }
      }
#endif
      if (blockIdx.x == 0 && tid == 0) a.cell_lb[Ng] = Ng - 1;
      grid.sync();
    }
    if (!SH && !use_win && use_tau && a.coarse_log2 >= 0) {        // the tiles' last thresholds into shared memory
      for (int k = tid; k < (1 << a.coarse_log2); k += NT) s_win[k] = __ldcg(a.tau_tile + k);
      __syncthreads();
    }
    if (!use_win) return;
    // Sorted positions: every answer of chunk q lies in [lower_bound(first slot of q), lower_bound(first slot of q+1)].
    // One bisection per chunk, all of them in flight together (one latency per step instead of one per chunk); on one
    // GPU its upper levels run on the table of the tiles' last thresholds in shared memory (11 dependent loads, not 24).
    const bool two = !SH && a.coarse_log2 >= 0;
    if (two) {
      for (int k = tid; k < (1 << a.coarse_log2); k += NT) s_win[k] = __ldcg(a.tau_tile + k);
      __syncthreads();
    }
    // Sharded filter: BRACKET FIRST.  Every rank holds the prefix of all tiles (Etile, phase C1), so the tiles between which
    // this CTA's ancestors lie are found without touching a peer; two probes of tau confirm the bracket [blo, bhi) (tau equals
    // the cdf only up to its last bits); the part of it that lives on other ranks is copied here in one coalesced sweep, and
    // the per-chunk bisections below run on local data.  (Before: 24 dependent NVLink round trips per bisection, ~65 us
    // in the CTAs whose ancestors are remote -- profiles/r2_k_shard_trace_g2.txt.)
    if constexpr (SH) {
      const int ntg = a.ntiles_g - 1;                  // tiles of the whole filter; Etile[ntg] = total
      const bool bracket = nchunk_valid > 0 && ntg <= CW * WINCAP && nwb <= PF_FUSED_MAXWB;
      if (bracket) {
        for (int k = tid; k < ntg; k += NT) s_win[k] = __ldcg(a.Etile + k);
        __syncthreads();
      }
      if (tid < 2 * PF_FUSED_MAXWB) {
        const int b = tid >> 1, up = tid & 1;
        int v = up ? Ng : 0;
        if (bracket && b < nwb) {
          const int slot = a.goff + (!strided ? (up ? min(cta_base + 128 * nchunk_cta, N) - 1 : cta_base)
                                              : (up ? pbase(b * wb + wb - 1) + 127 : pbase(b * wb)));
          const float r = position(slot, res_u0);
          // lower end: the tiles that END below r (Etile[k + 1] < r) hold no answer;  upper end: nor do the tiles that START above r
          int lo_ = 0, n_ = ntg;
          while (n_ > 0) {
            const int h = n_ >> 1, k = lo_ + h;
            const float e = up ? s_win[k] : ((k + 1 < ntg) ? s_win[k + 1] : res_total);
            const bool below = up ? (e <= r) : (e < r);
            if (below) { lo_ += h + 1; n_ -= h + 1; } else n_ = h;
          }
          // confirm against tau itself, which equals the cdf only up to its last bits: widen by a tile, then give up (NaN: give up)
          if (!up) {
            v = min(lo_ * TILE, Ng);
            if (v > 0 && (r <= tauG(v - 1))) v = max(v - TILE, 0);
            if (v > 0 && (r <= tauG(v - 1))) v = 0;
            if (!(r == r)) v = 0;
          } else {
            v = (lo_ > 0) ? min(lo_ * TILE, Ng) : Ng;
            if (v < Ng && !(r <= tauG(v - 1))) v = min(v + TILE, Ng);
            if (v < Ng && !(r <= tauG(v - 1))) v = Ng;
            if (!(r == r)) v = Ng;
          }
        }
        s_br[tid] = v;
      }
      __syncthreads();
      if (tid < PF_FUSED_MAXWB) {
        int on = 0;
        if (bracket && tid < nwb && a.stg_tau && tid < a.stg_blocks) {
          const int blo = s_br[2 * tid], bhi = s_br[2 * tid + 1];
          const bool remote = bhi > blo && (((blo >> a.log2ns) != a.rank) || (((bhi - 1) >> a.log2ns) != a.rank));
          on = (remote && bhi - blo <= a.stg_cap) ? 1 : 0;
        }
        s_stg[tid] = on;
      }
      __syncthreads();
      if (bracket && a.stg_tau) {
        // 16-byte peer loads, eight in flight per thread (128 KB per CTA and NVLink round trip), all blocks in one sweep; the
        // brackets and the ranks' boundaries are multiples of the tile length, so a vector never straddles two ranks
        float4* dt4 = reinterpret_cast<float4*>(a.stg_tau + (size_t)blockIdx.x * a.stg_blocks * a.stg_cap);
        const int cap4 = a.stg_cap >> 2, items = min(nwb, a.stg_blocks) * cap4;
        for (int i0 = 0; i0 < items; i0 += 8 * NT) {
          float4 tv[8];
          bool rm[8];
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            const int i = min(i0 + tid + m * NT, items - 1), b = i / cap4, o = i - b * cap4;
            const int g = s_br[2 * b] + 4 * o;
            const bool in = (i0 + tid + m * NT < items) && s_stg[b] && g < s_br[2 * b + 1];
            const int r = in ? (g >> a.log2ns) : a.rank;
            rm[m] = in && r != a.rank;
            tv[m] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (rm[m]) tv[m] = __ldcg(reinterpret_cast<const float4*>(a.tau_peers[r] + (g & ns_mask)));
          }
#pragma unroll
          for (int m = 0; m < 8; ++m)
            if (rm[m]) dt4[i0 + tid + m * NT] = tv[m];
        }
        __syncthreads();
      }
    }
    // one lower bound per chunk boundary: per window block, its wb chunk starts and its last slot
    const int nbound = strided ? nwb * (wb + 1) : nchunk_valid + 1;
    for (int kb = tid; kb < nbound; kb += NT) {
      int slot, blk = 0;
      if (!strided) slot = (kb < nchunk_valid) ? cta_base + 128 * kb : min(cta_base + 128 * nchunk_cta, N) - 1;
      else {
        blk = kb / (wb + 1);
        const int e = kb - blk * (wb + 1);
        slot = (e < wb) ? pbase(blk * wb + e) : pbase(blk * wb + wb - 1) + 127;
      }
      int lb = 0;
      if (nchunk_valid > 0) {
        const float r = position(a.goff + slot, res_u0);
        if (two) lb = lower_bound_two_level(s_win, a.coarse_log2, a.E, r);
        else if (!SH) lb = lower_bound_tau_pow2(a.E, a.log2ng, r);
        else {                                        // lower bound inside the bracket (tau is monotone: any search order)
          set_stage(blk);
          int pos = s_br[2 * blk], len = s_br[2 * blk + 1] - pos;
          while (len > 0) {
            const int half = len >> 1;
            if (!(r <= tauG(pos + half))) { pos += half + 1; len -= half + 1; } else len = half;
          }
          lb = pos;
        }
      }
      s_lo[kb] = lb;
    }
    __syncthreads();
  };

  // Ancestors of the 4 slots base..base+3 of the chunk starting at cbase.  Warp-synchronous (all lanes call it).
  // one global lower bound (any rank's shard) -- used by the rare fallbacks and by unsorted (multinomial) positions
  auto lb_global = [&](float r) -> int {
    if (!SH) return lower_bound_tau_pow2(a.E, a.log2ng, r);
    unsigned low = 0;
    for (unsigned sft = 1u << (a.log2ng - 1); sft != 0; sft >>= 1)
      if (!(r <= tauG((int)(low + sft)))) low += sft;
    return (int)(low + ((r <= tauG((int)low)) ? 0u : 1u));
  };

  // Ancestors (GLOBAL particle indices) of the 4 slots base..base+3 (local) of chunk q.  Warp-synchronous.
  auto ancestors4 = [&](int ts, int q, int base, int nv, int anc[4], int& xs_lo) {
    xs_lo = -1;                                        // >= 0: the chunk's ancestors x[xs_lo ...] are staged in shared memory
    float u[4] = {res_u0, res_u0, res_u0, res_u0};
    if (a.resampler != 1 && nv > 0) {
      if (a.res_u) {
        const float* p = a.res_u + (ptrdiff_t)ts * a.B * N;
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = (k < nv) ? p[base + k] : 0.0f;
      } else {
        philox_quad_uniforms(a.seed, BF_STREAM_RESAMPLE, (uint32_t)(ts + a.t_offset), (uint32_t)a.trial, (uint32_t)((a.goff + base) >> 2), u);
      }
    }
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = position(a.goff + base + k, u[k]);
    if (use_win) {
      const int lo = s_lo[sidx(q)], wn = s_lo[sidx(q) + 1] - lo;    // warp-uniform
      if (SH) set_stage(strided ? q / wb : 0);
#ifndef PF_FUSED_COUNTING
#define PF_FUSED_COUNTING 1
#endif
#if PF_FUSED_COUNTING
      if (wn <= 32 * 128) {
        count_window(lo, wn, a.goff + pbase(q), res_u0, anc);
#else
      if (wn <= WINCAP) {
        float* win = s_win + warp * WINCAP;
        for (int i0 = 0; i0 < wn; i0 += 128) {          // four independent loads in flight per lane
          float tv[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) { const int i = i0 + lane + 32 * m; tv[m] = (i < wn) ? tauG(lo + i) : 0.0f; }
#pragma unroll
          for (int m = 0; m < 4; ++m) { const int i = i0 + lane + 32 * m; if (i < wn) win[i] = tv[m]; }
        }
        __syncwarp();
        // a lane's four answers lie between its first answer and the next lane's first answer
        int p = lower_bound_window(win, 0, wn, r[0]);
        int pn = __shfl_down_sync(BF_FULL, p, 1);
        if (lane == 31) pn = wn;
        anc[0] = min(lo + p, Ng - 1);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          p = lower_bound_window(win, p, pn, r[k]);
          anc[k] = min(lo + p, Ng - 1);
        }
#endif
        // Sharded filter, ancestors (partly) on another rank: x[lo .. lo+wn] is a contiguous range -- fetch it with
        // coalesced 512-byte peer reads into shared memory instead of one 16-byte NVLink round trip per particle.
        if (SH && a.xstage && wn + 1 <= PF_FUSED_XCAP && (((lo >> a.log2ns) != a.rank) || ((min(lo + wn, Ng - 1) >> a.log2ns) != a.rank))) {
          if (x_pre) { xs_lo = lo; }                   // already on its way (xprefetch), same window, same test
          else {
          float* xs = s_xst + (size_t)(2 * warp + x_buf) * PF_FUSED_XCAP * DX;
          for (int i0 = 0; i0 <= wn; i0 += 128) {       // four peer loads in flight per lane (an NVLink round trip is ~3 us)
            float xv[4][DX];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int g = min(lo + min(i0 + lane + 32 * m, wn), Ng - 1);
              load_state_cg<DX>(xp_stage[g >> a.log2ns], (size_t)(g & ns_mask), xv[m]);
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int i = i0 + lane + 32 * m;
              if (i <= wn) {
#pragma unroll
                for (int c = 0; c < DX; ++c) xs[(size_t)i * DX + c] = xv[m][c];
              }
            }
          }
          xs_lo = lo;
          }
        }
        __syncwarp();                                  // the window is overwritten by this warp's next chunk
      } else if (nv > 0) {                             // window larger than the cap (very uneven weights): rare
#pragma unroll
        for (int k = 0; k < 4; ++k) anc[k] = min(lb_global(r[k]), Ng - 1);
      }
      return;
    }
    if (nv <= 0) return;
    if (use_tau) {
      if (use_cells) {
        // w = 1 - u as position() rounds it; cells only for w in [0, 1] (anything else -- a caller's own "uniforms" -- takes
        // the general search below)
        float wq[4]; int c[4], lo[4], hi[4];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          wq[k] = __fmul_rn(__fsub_rn(1.0f, u[k]), (float)Ng);          // exact (power of two)
          ok = ok && (wq[k] >= 0.0f) && (wq[k] <= (float)Ng);
          c[k] = min(max((int)wq[k], 0), Ng - 1);
        }
        if (ok) {
#if defined(BF_EXP_MN_NOLOOKUP)
#pragma unroll
          for (int k = 0; k < 4; ++k) anc[k] = a.goff + base + k;
          return;
#elif defined(BF_EXP_MN_XONLY)
#pragma unroll
          for (int k = 0; k < 4; ++k) anc[k] = c[k];
          return;
#endif
#pragma unroll
          for (int k = 0; k < 4; ++k) {                  // (r == the cell's lower edge: its entry is the answer, the next one is not read)
            lo[k] = a.cell_lb[c[k]];
            hi[k] = (wq[k] == (float)c[k]) ? 0 : a.cell_lb[c[k] + 1];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int pos = lo[k], len = (wq[k] == (float)c[k]) ? 0 : hi[k] - lo[k];
            while (len > 0) {
              const int half = len >> 1;
              if (!(r[k] <= a.E[pos + half])) { pos += half + 1; len -= half + 1; } else len = half;
            }
            anc[k] = min(pos, Ng - 1);
          }
          return;
        }
      }
      if (!SH && a.coarse_log2 >= 0) lower_bound4_two_level(s_win, a.coarse_log2, a.E, r, Ng - 1, anc);
      else if (!SH) bisect4_left_exclusive_pow2(a.E - 1, a.log2ng, r, anc);       // ex[j+1] == tau[j]
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) anc[k] = min(lb_global(r[k]), Ng - 1);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) anc[k] = fused_bisect_left_exclusive(a.E, N, levels, r[k]);
    }
  };

  // particle `g` (GLOBAL index) of the state the step reads: own shard or a peer's
  auto load_prev = [&](int g, const float* xp_local, const float* const* xp_peers, float* x) {
    if (!SH) load_state_cg<DX>(xp_local, (size_t)g, x);
    else load_state_cg<DX>(xp_peers[g >> a.log2ns], (size_t)(g & ns_mask), x);
  };

  const bool timer = (blockIdx.x == 0 && tid == 0 && a.phase_ns != nullptr);
  unsigned long long tmark = timer ? globaltimer_ns() : 0ull;
  auto lap = [&](int phase) {
    if (timer) { const unsigned long long now = globaltimer_ns(); a.phase_ns[phase] += now - tmark; tmark = now; }
  };

#ifdef BF_EXP_TRACE   // experiment: per-CTA arrival / release times at the four grid syncs of step T/2 -> phase_ns[8 + 16 b + k], SM id at k = 8
  auto trace = [&](int t, int k) {
    if (tid == 0 && a.phase_ns != nullptr && t == T / 2) {
      a.phase_ns[8 + 16 * blockIdx.x + k] = globaltimer_ns();
      if (k == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); a.phase_ns[8 + 16 * blockIdx.x + 8] = sm; }
    }
  };
#else
  auto trace = [&](int, int) {};
#endif

  // Fold `nb` roots as the balanced binary tree over the index: a thread takes R consecutive roots (R = 1 up to 2^27
  // particles), a warp its 32 threads by a butterfly of merges (adjacent lanes first), warp 0 the warps' results the same way.
  // The result is in thread 0's acc.  Whoever folds gets the same bits.  All threads of the CTA call it.
  auto fold_roots = [&](OnlineAcc<DX>& acc, const PfPartial* roots, int nb) {
    int P2 = 1;
    while (P2 < nb) P2 <<= 1;
    const int R = max(1, P2 / NT);
    acc.init();
    if (R == 1) { if (tid < nb) acc.load(roots + tid); }
    else if (R == 2) {
      if (2 * tid < nb) acc.load(roots + 2 * tid);
      if (2 * tid + 1 < nb) { OnlineAcc<DX> o; o.load(roots + 2 * tid + 1); acc.merge_acc(o); }
    } else {
      PfPartial tmp;
      fold_root_range<DX>(roots, tid * R, R, nb, &tmp);
      acc.load_plain(&tmp);
    }
    acc.warp_tree_reduce();
    PfPartial* tr = reinterpret_cast<PfPartial*>(s_win);      // [CW], the tau-window area, idle here
    __syncthreads();
    if (lane == 0) acc.store(tr + warp);
    __syncthreads();
    if (warp == 0) {
      acc.init();
      if (lane < CW) acc.load_plain(tr + lane);
      acc.warp_tree_reduce();
    }
  };
  // thread 0's acc (the fold of the whole filter) -> the step's statistics (s_stats) and the per-step outputs
  auto finish_statistics = [&](OnlineAcc<DX>& acc, int t, int pt, bool outputs_by_cta0) {
    if (tid == 0) {
      StepStats st;
      st.mx = acc.m; st.Z = acc.s;
      st.loglik = acc.m + logf(acc.s);                       // log p(y_t | y_1:t-1)
      st.ess = acc.s * acc.s / acc.s2;                       // 1 / sum (e/Z)^2   (inference.py:1356)
      st.flag = (st.ess < a.ess_threshold * (float)Ng) ? 1 : 0;
      st.w_uniform = st.flag;
      st.pad0 = st.pad1 = 0;
#pragma unroll
      for (int k = 0; k < BF_MAXD; ++k) st.mean[k] = (k < DX) ? acc.sx[k] / acc.s : 0.0f;
      s_stats = st;
      if (blockIdx.x == 0 && !all_phases) a.carry_stats[(pt + 1) & 1] = st;
      if ((blockIdx.x == 0 || !outputs_by_cta0) && a.rank == 0) {
        if (a.out_loglik) a.out_loglik[t] = st.loglik;
        if (a.out_ess) a.out_ess[t] = st.ess;
        if (a.out_resampled) a.out_resampled[t] = st.flag;
        if (a.out_mean) {
#pragma unroll
          for (int k = 0; k < DX; ++k) a.out_mean[(size_t)t * DX + k] = st.mean[k];
        }
      }
    }
    __syncthreads();
  };

  // single-phase launches (sharded filter): one step per launch, t_offset is its absolute index; the state that
  // crosses launches lives in two StepStats slots, read at [pt & 1] and written (phase B) at [(pt + 1) & 1]
  if (!all_phases) {
    const StepStats* sin = a.carry_stats + (((a.phases & PF_PH_C) ? a.t_offset + 1 : a.t_offset) & 1);
    w_uni = sin->w_uniform != 0; gather = sin->flag != 0;
    if (a.t_offset > 0) x_prev = (a.t_offset & 1) ? a.xbuf0 : a.xbuf1;
  }
  for (int t = 0; t < T; ++t) {
    const int pt = all_phases ? t : a.t_offset;        // ping-pong parity and "is there a previous step"
    float* x_cur = (pt & 1) ? a.xbuf1 : a.xbuf0;
    const float* const* xp_peers = (pt & 1) ? a.xa_peers : a.xb_peers;
    const float u_in = a.inputs ? a.inputs[t] : 0.0f;
    float y[DY];
#pragma unroll
    for (int k = 0; k < DY; ++k) y[k] = a.emissions[(size_t)t * DY + k];
    OnlineAcc<DX> acc;

    // ------------------------------------------------------------------ phase A
    if (a.phases & PF_PH_A) {
    trace(t, 9);
    xp_stage = xp_peers;
    if (gather) begin_gather(t - 1);
    trace(t, 10);
    // Warps take UNITS of 4 consecutive chunks (512 particles): a lane's 16 particles are accumulated in index order and the
    // unit is reduced at once, so the unit partial is a function of the data alone (see PfFusedArgs::upart).
    // (sharded) rows of chunk q's ancestor window -> buffer b of this warp, asynchronously; always commits one group
    auto xprefetch = [&](int q, int b) {
      if constexpr (XPIPE) {
        if (q >= 0 && q < nchunk_valid) {
          const int lo = s_lo[sidx(q)], wn = s_lo[sidx(q) + 1] - lo;
          if (wn + 1 <= PF_FUSED_XCAP && (((lo >> a.log2ns) != a.rank) || ((min(lo + wn, Ng - 1) >> a.log2ns) != a.rank))) {
            float* xs = s_xst + (size_t)(2 * warp + b) * PF_FUSED_XCAP * DX;
            for (int i = lane; i <= wn; i += 32) {
              const int g = min(lo + i, Ng - 1);
              const float* src = xp_stage[g >> a.log2ns] + (size_t)(g & ns_mask) * DX;
              const unsigned dst = (unsigned)__cvta_generic_to_shared(xs + (size_t)i * DX);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
      }
    };
    const int nunit_cta = nchunk_cta / 4;
    const bool xpipe = XPIPE && gather && use_win && a.xstage != 0 && a.xpipe != 0;
    int un = next_chunk(0);
    x_pre = xpipe; x_buf = 0;
    if (xpipe) xprefetch(un < nunit_cta ? 4 * un : -1, 0);
    while (un < nunit_cta) {
    const int un_next = xpipe ? next_chunk(0) : 0;      // taken early: its first chunk is prefetched during this unit's last
    const int lu = pbase(4 * un) / 512;                // unit index within this rank
    if (lu >= a.nunits) break;                         // tile padding past N
    acc.init();
#pragma unroll 1                                        // the body is ~30 KB of SASS: four copies would not fit the instruction cache
    for (int c4 = 0; c4 < 4; ++c4) {
      const int q = 4 * un + c4;
      if (q >= nchunk_valid) break;
      const int base = pbase(q) + lane * 4;
      const int nv = max(0, min(4, N - base));
      int anc[4] = {a.goff + base, a.goff + base + 1, a.goff + base + 2, a.goff + base + 3};     // global indices
      int xs_lo = -1;
      xp_stage = xp_peers;
      if (xpipe) {                                     // next chunk's rows on their way; this chunk's have landed
        __syncwarp();                                  // (every lane has read the rows of the chunk before last out of that buffer)
        xprefetch((c4 < 3) ? q + 1 : (un_next < nunit_cta ? 4 * un_next : -1), x_buf ^ 1);
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
      }
#ifndef BF_EXP_NOSEARCH
      if (gather) ancestors4(t - 1, q, base, nv, anc, xs_lo);
#endif
      const int xb_cur = x_buf;
      if (xpipe) x_buf ^= 1;                           // (the rows of this chunk are read below through xb_cur)
      if (nv <= 0) continue;
      if (pt > 0 && a.out_ancestors) {
        int* o = a.out_ancestors + ((ptrdiff_t)t - 1) * N + base;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < nv) o[k] = anc[k];
      }
      constexpr bool RNG_PER_PARTICLE = (DQ % 4 == 0);
      float z[RNG_PER_PARTICLE ? 1 : 4][DQ];
      if (!a.noise && !RNG_PER_PARTICLE)
        philox_quad_normals<DQ>(a.seed, BF_STREAM_NOISE, (uint32_t)(t + a.t_offset), (uint32_t)a.trial, (uint32_t)((a.goff + base) >> 2),
                                reinterpret_cast<float(*)[DQ]>(z));
      // The four particles of a lane: their states are fetched together (four gathers in flight), then propagated two at
      // a time in a loop that is NOT unrolled further -- the body (Philox colouring, f, log-likelihood, statistics) is
      // ~600 instructions per particle and the phase is instruction-fetch-bound when all four copies are laid out.
      // The second pair moves into the first pair's registers between the two turns.
      float xs[4][DX];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k < nv) {
          if (SH && xs_lo >= 0) {
            const float* xsrc = s_xst + ((size_t)(2 * warp + xb_cur) * PF_FUSED_XCAP + (anc[k] - xs_lo)) * DX;
#pragma unroll
            for (int c = 0; c < DX; ++c) xs[k][c] = xsrc[c];
          } else if (gather) load_prev(anc[k], x_prev, xp_peers, xs[k]);
          else load_state_cg<DX>(x_prev, (size_t)(base + k), xs[k]);
        } else {
#pragma unroll
          for (int c = 0; c < DX; ++c) xs[k][c] = 0.0f;
        }
      }
      float llv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#ifndef PF_FUSED_PU
#define PF_FUSED_PU 2                        // particles per turn of the propagate loop
#endif
      constexpr int PU = PF_FUSED_PU;
#pragma unroll 1
      for (int h = 0; h < 4 / PU; ++h) {
        float lh[PU], wv[PU], xt[PU][DX];
#pragma unroll
        for (int kk = 0; kk < PU; ++kk) {
          const int k = PU * h + kk;
          lh[kk] = -INFINITY; wv[kk] = 0.0f;           // a slot past N: weight exp(-inf) * 0, never stored
#pragma unroll
          for (int c = 0; c < DX; ++c) xt[kk][c] = 0.0f;
          if (LEAN || k < nv) {
            const int i = base + k;
            float qn[DQ];
            const float* x = xs[kk];
            if (pt > 0 && a.out_particles) store_state<DX>(a.out_particles + ((ptrdiff_t)t - 1) * N * DX, (size_t)i, x);
            if (a.noise) {
              const float* np_ = a.noise + (size_t)t * a.B * N * DQ + (size_t)i * DQ;
#pragma unroll
              for (int c = 0; c < DQ; ++c) qn[c] = np_[c];
            } else {
              if (RNG_PER_PARTICLE)
                philox_particle_normals<DQ>(a.seed, BF_STREAM_NOISE, (uint32_t)(t + a.t_offset), (uint32_t)a.trial,
                                            (uint32_t)((a.goff + base) >> 2), (uint32_t)k, z[0]);
              const float* zk = z[RNG_PER_PARTICLE ? 0 : kk];
#pragma unroll
              for (int r = 0; r < DQ; ++r) {             // q = q0 + L_Q z  (models.py:83, TFP MVN sample)
                float s = mc.q0[r];
#pragma unroll
                for (int c = 0; c <= r; ++c) s = fmaf(mc.LQ[r * DQ + c], zk[c], s);
                qn[r] = s;
              }
            }
            M::f(x, qn, u_in, mc, xt[kk]);
            store_state_cs<DX>(x_cur, (size_t)i, xt[kk]);
            if (a.out_propagated) store_state<DX>(a.out_propagated + (size_t)t * N * DX, (size_t)i, xt[kk]);
#ifdef BF_EXP_NOLL
            lh[kk] = -0.5f * xt[kk][0] * xt[kk][0];
#else
            lh[kk] = emission_log_prob<M>(xt[kk], y, u_in, mc);
#endif
            wv[kk] = w_uni ? invN : a.W[i];
          }
        }
        acc.template add_group<PU>(lh, wv, xt);
        // shift everything by PU particles: results in at the top, the next particles' state and noise down to slot 0
#pragma unroll
        for (int k = 0; k < 4 - PU; ++k) llv[k] = llv[k + PU];
#pragma unroll
        for (int kk = 0; kk < PU; ++kk) llv[4 - PU + kk] = lh[kk];
#pragma unroll
        for (int k = 0; k < 4 - PU; ++k) {
#pragma unroll
          for (int c = 0; c < DX; ++c) xs[k][c] = xs[k + PU][c];
          if (!RNG_PER_PARTICLE) {
#pragma unroll
            for (int c = 0; c < DQ; ++c) z[RNG_PER_PARTICLE ? 0 : k][c] = z[RNG_PER_PARTICLE ? 0 : k + PU][c];
          }
        }
      }
      *reinterpret_cast<float4*>(a.LL + base) = make_float4(llv[0], llv[1], llv[2], llv[3]);
    }
    // unit partial -> global memory (plain stores: it is read by this CTA only, after the barrier below)
#ifndef BF_EXP_NOUNITRED
    acc.unit_reduce();
#endif
    if (lane == 0) acc.store(a.upart + lu);
    un = xpipe ? un_next : next_chunk(0);
    }  // units
    if (xpipe) { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
    x_pre = false; x_buf = 0;
    trace(t, 11);
    if (tid == 0) { s_ctr[1] = 0; s_ctr[2] = 0; }      // queues of phases B and C2 (nobody is in them now)
#ifndef BF_DYNAMIC_CHUNKS
    turn[1] = 0; turn[2] = 0;
#endif
    __syncthreads();
    {
      // block roots of this CTA: a warp folds the 4 bt units of a block (lane = unit, butterfly of merges, adjacent first)
      const int ub = 4 * a.bt;
      const int nblk_cta = (ntile_cta + a.bt - 1) / a.bt;
      for (int bl = warp; bl < nblk_cta; bl += CW) {
        const int gb = strided ? bl * (int)gridDim.x + (int)blockIdx.x : tile0 / a.bt + bl;      // block of the rank
        const int lu0 = gb * ub;
        OnlineAcc<DX> v;
        v.init();
        if (lane < ub && lu0 + lane < a.nunits) v.load(a.upart + lu0 + lane);
        v.warp_tree_reduce();
        if (lane == 0) v.store(a.lroot + gb);
      }
      __syncthreads();
      // the LAST CTA of this rank out of the phase (ticket) folds the rank's block roots once, before the grid leaves the
      // barrier; the others fetch the result (every CTA folding them itself was an 11 us L2 hot spot per step)
      if (tid == 0) {
        __threadfence();
        const unsigned tkc = atomicAdd(a.fold_ctr, 1u);
        s_last = ((tkc + 1u) % (unsigned)G == 0u) ? 1 : 0;
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        fold_roots(acc, a.lroot, a.nroots);
        if (!SH) {
          finish_statistics(acc, t, pt, false);
          if (tid == 0) {
            *a.fold_stats = s_stats;
            __threadfence();
          }
        } else if (tid == 0) {                         // sharded: this rank's root, to itself and to every peer
          const size_t ridx = (size_t)(pt & 1) * a.world + a.rank;
          acc.store(a.broot + ridx);
          if (p2p) {
            for (int r = 0; r < a.world; ++r) {
              if (r == a.rank) continue;
              acc.store(reinterpret_cast<PfPartial*>(a.xchg_peers[r] + a.off_root) + ridx);
            }
            __threadfence_system();
          }
          __threadfence();
        }
      }
    }
    }  // phase A
    trace(t, 0);
    if (all_phases) grid.sync();
    trace(t, 1);
    if (p2p) xbarrier(0, t);
    trace(t, 12);
    lap(0);
    if (!(a.phases & (PF_PH_B | PF_PH_C))) break;

    // ------------------------------------------------------------------ phase B
    if (tid == 0) s_ctr[0] = 0;                        // queue of the next phase A / the final gather
#ifndef BF_DYNAMIC_CHUNKS
    turn[0] = 0;
#endif
    if (!(a.phases & PF_PH_B)) {                       // C-only launch: phase B left the step's statistics here
      if (tid == 0) s_stats = a.carry_stats[(pt + 1) & 1];
      __syncthreads();
    } else {
    if (SH || !all_phases) { fold_roots(acc, a.broot + (size_t)(pt & 1) * a.world, a.world); finish_statistics(acc, t, pt, true); }
    else if (tid == 0) {                               // one-GPU run: the last CTA out of phase A folded them (below)
      const int4* src = reinterpret_cast<const int4*>(a.fold_stats);
      int4* dst = reinterpret_cast<int4*>(&s_stats);
#pragma unroll
      for (int k = 0; k < (int)(sizeof(StepStats) / 16); ++k) dst[k] = __ldcg(src + k);
    }
    __syncthreads();
    }  // statistics
    const float mx = s_stats.mx, rZ = 1.0f / s_stats.Z;
    const bool flag = s_stats.flag != 0;
    if (a.phases & PF_PH_B) {

    // normalised weights (exp(ll - max) * w_prev) / sum (inference.py:1350-1353), four chunks per turn so that four
    // 512-byte loads per warp are in flight (one at a time left this phase latency-bound)
    for (int qb = next_chunks4(1); qb < nchunk_cta; qb = next_chunks4(1)) {
      float4 l4[4], wp4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int base = pbase(qb + j) + lane * 4;
        l4[j] = *reinterpret_cast<const float4*>(a.LL + base);
        wp4[j] = make_float4(invN, invN, invN, invN);
        if (!w_uni && w_vec && base + 3 < N) wp4[j] = *reinterpret_cast<const float4*>(a.W + base);
        else if (!w_uni) {
          wp4[j].x = (base < N) ? a.W[base] : 0.0f; wp4[j].y = (base + 1 < N) ? a.W[base + 1] : 0.0f;
          wp4[j].z = (base + 2 < N) ? a.W[base + 2] : 0.0f; wp4[j].w = (base + 3 < N) ? a.W[base + 3] : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = qb + j, base = pbase(q) + lane * 4;
        const float ll[4] = {l4[j].x, l4[j].y, l4[j].z, l4[j].w}, wp[4] = {wp4[j].x, wp4[j].y, wp4[j].z, wp4[j].w};
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (base + k < N) ? norm_weight_rcp(ll[k], wp[k], mx, rZ) : 0.0f;
        if (a.out_norm_weights) {
#pragma unroll
          for (int k = 0; k < 4; ++k) if (base + k < N) a.out_norm_weights[(size_t)t * N + base + k] = v[k];
        }
        if (!flag) {                     // identity branch of lax.cond (inference.py:1357): carry the weights
#pragma unroll
          for (int k = 0; k < 4; ++k) if (base + k < N) {
            a.W[base + k] = v[k];
            if (a.out_weights) a.out_weights[(size_t)t * N + base + k] = v[k];
          }
        } else {
          if (a.out_weights) {           // utils.py:212: weights = 1/N after resampling
#pragma unroll
            for (int k = 0; k < 4; ++k) if (base + k < N) a.out_weights[(size_t)t * N + base + k] = invN;
          }
          *reinterpret_cast<float4*>(a.E + base) = make_float4(v[0], v[1], v[2], v[3]);   // for phase C2
          const float cs = chunk_tree_sum(v);
          if (lane == 0) { s_wsum[q] = cs; s_v0[q] = v[0]; }
        }
      }
    }
    }  // phase B weights
    x_prev = x_cur;
    if (!flag) {                         // grid-uniform branch: no grid sync, but chunks change hands between warps
      __syncthreads();                   // W of this step is complete before any warp of the CTA starts the next step
      w_uni = false; gather = false; lap(1); continue;
    }
    if (a.phases & PF_PH_B) {
    __syncthreads();
    for (int i = tid; i < ntile_cta; i += NT) {       // tile sums: the top log2(TW) levels of the tile's tree
      float y16[TW];
#pragma unroll
      for (int k = 0; k < TW; ++k) y16[k] = s_wsum[i * TW + k];
#pragma unroll
      for (int w = TW / 2; w >= 1; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) y16[k] = __fadd_rn(y16[2 * k], y16[2 * k + 1]);
      const int gi = a.tile_off + tile_of(i);
      a.tilesum[gi] = y16[0];
      a.V[gi] = s_v0[i * TW];
      if (p2p) {
        for (int r = 0; r < a.world; ++r) {
          if (r == a.rank) continue;
          reinterpret_cast<float*>(a.xchg_peers[r] + a.off_ts)[gi] = y16[0];
          reinterpret_cast<float*>(a.xchg_peers[r] + a.off_V)[gi] = s_v0[i * TW];
        }
        __threadfence_system();
      }
    }
    }  // phase B tile sums
    trace(t, 2);
    if (all_phases) grid.sync();
    trace(t, 3);
    if (p2p) xbarrier(1, t);
    trace(t, 13);
    lap(1);
    if (!(a.phases & PF_PH_C)) break;

    if (!(a.phases & PF_PH_B)) {
      // C-only launch (nccl transport): phase B ran in ANOTHER launch, whose shared memory is gone -- rebuild the chunk tree
      // sums and first weights from the normalised weights it left in E (same additions, same bits)
      for (int q = warp; q < nchunk_cta; q += CW) {
        const float4 w4 = *reinterpret_cast<const float4*>(a.E + pbase(q) + lane * 4);
        const float v[4] = {w4.x, w4.y, w4.z, w4.w};
        const float cs = chunk_tree_sum(v);
        if (lane == 0) { s_wsum[q] = cs; s_v0[q] = v[0]; }
      }
      __syncthreads();
    }
    // ------------------------------------------------------------------ phase C1 (CTA 0; every rank of a sharded filter)
    if (c1_all) {
      // EVERY CTA scans the tile sums itself, into its own shared memory: the same tree, the same bits, and the grid does not
      // meet again between C1 and C2 (one CTA scanning while 146 wait, then a barrier: ~10 us per step).  CTA 0 keeps the
      // global copy current (the sharded filter's ancestor brackets and E[Ng] read it in the next step).
      for (int i = tid; i < a.ntiles_g; i += NT) s_win[i] = __ldcg(a.tilesum + i);
      __syncthreads();
      staged_tile_scan<CW>(s_win, s_et, a.ntiles_g, s_win + ((a.ntiles_g + 3) & ~3));
      __syncthreads();
      if (blockIdx.x == 0) for (int i = tid; i <= a.ntiles_g; i += NT) a.Etile[i] = s_et[i];
      trace(t, 4);
      trace(t, 5);
    } else {
    if (blockIdx.x == 0) {
      if (a.stage_ts) {                                // all the loads in flight at once: one L2 latency instead of ten
        for (int i = tid; i < a.ntiles_g; i += NT) s_win[i] = __ldcg(a.tilesum + i);
        __syncthreads();
        staged_tile_scan<CW>(s_win, a.Etile, a.ntiles_g, s_win + ((a.ntiles_g + 3) & ~3));
      } else {
        cta_exclusive_scan<CW>(a.tilesum, a.Etile, a.ntiles_g, s_win);
      }
    }
    trace(t, 4);
    grid.sync();
    trace(t, 5);
    }
    lap(2);

    // ------------------------------------------------------------------ phase C2
    // The tile-level part of the scan and of tau (prefix and clamp interval of each chunk) is the same for all 32
    // lanes of a chunk's warp: one thread per TILE computes it for the tile's 16 chunks (tile_levels) into the
    // tau-window area of shared memory, idle in this phase.  Phase C2 is bound by the ALU pipe (FMNMX / FSEL / ISETP
    // of the clamps), so this takes a fifth of its instructions away.  Very large tiles_per_cta: per-lane path.
    const bool c2_pre = use_tau && (3 * TW + 1) * tpc <= CW * WINCAP;
    float* s_cE = s_win;                               // [(TW+1)*tpc]
    float* s_clo = s_cE + (TW + 1) * tpc;              // [TW*tpc]
    float* s_chi = s_clo + TW * tpc;                   // [TW*tpc]
    if (use_tau) {
      // clamp interval of each owned tile from the bisection levels above it (half-widths N/2 .. TILE); the probe
      // value of a node whose middle is the first boundary of tile m is cdf[m*TILE] = Etile[m] + w[m*TILE]
      for (int i = tid; i < ntile_cta; i += NT) {
        const int tile = a.tile_off + tile_of(i);
        const int jt = tile * TILE;                              // global boundary index
        float lo = -INFINITY, hi = INFINITY;
        if (jt < Ng) {
          // the probes do not depend on the clamps: fetch them all first (one L2 latency, not one per level)
          constexpr int MAXLV = 20;                              // N up to 2^31 / tile of 2^11
          float cv[MAXLV];
#pragma unroll
          for (int l = 0; l < MAXLV; ++l) {
            const int s = (Ng >> 1) >> l;
            cv[l] = 0.0f;
            if (s >= TILE) {
              const int m = ((jt & ~(2 * s - 1)) + s) / TILE;
              cv[l] = __fadd_rn(etile(m), __ldcg(&a.V[m]));
            }
          }
#pragma unroll
          for (int l = 0; l < MAXLV; ++l) {
            const int s = (Ng >> 1) >> l;
            if (s >= TILE) {
              const int mid = (jt & ~(2 * s - 1)) + s;
              const float c = clampf(cv[l], lo, hi);
              if (jt < mid) hi = c; else lo = c;
            }
          }
        }
        s_meta_f[2 * i] = lo; s_meta_f[2 * i + 1] = hi;
        if (c2_pre)
          tile_levels<TW>(s_wsum + i * TW, s_v0 + i * TW, etile(tile), etile(tile + 1), lo, hi, min(TILE, Ng),
                          s_cE + i * (TW + 1), s_clo + i * TW, s_chi + i * TW);
      }
      __syncthreads();
    }
    for (int qb = next_chunks4(2); qb < nchunk_cta; qb = next_chunks4(2)) {
      float4 w4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)        // the weights phase B left here
        w4[j] = *reinterpret_cast<const float4*>(a.E + pbase(qb + j) + lane * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = qb + j, tl = q / TW, wpos = q % TW, tile = a.tile_off + tile_of(tl);
        const int base = pbase(q) + lane * 4;
        float e[4], tau[4];
        const float v[4] = {w4[j].x, w4[j].y, w4[j].z, w4[j].w};
        float4 out;
        if (c2_pre) {
          float s01, recv[5];
          chunk_upsweep(v, s01, recv);
          const float* ce = s_cE + tl * (TW + 1) + wpos;
          chunk_downsweep<true>(v, s01, recv, ce[0], ce[1], s_clo[q], s_chi[q], min(TILE, Ng), e, tau);
          out = make_float4(tau[0], tau[1], tau[2], tau[3]);
        } else if (use_tau) {
          chunk_scan<TW, true>(v, wpos, s_wsum + tl * TW, s_v0 + tl * TW, etile(tile), etile(tile + 1),
                                   s_meta_f[2 * tl], s_meta_f[2 * tl + 1], min(TILE, Ng), e, tau);
          out = make_float4(tau[0], tau[1], tau[2], tau[3]);
        } else {
          chunk_scan<TW, false>(v, wpos, s_wsum + tl * TW, s_v0 + tl * TW, etile(tile), 0.0f, 0.0f, 0.0f, 0, e, tau);
          out = make_float4(e[0], e[1], e[2], e[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) if (base + k == N) *a.total = e[k];          // E[N], the canonical total
        if (!use_tau || base < N) *reinterpret_cast<float4*>(a.E + base) = out;
        if (!SH && a.coarse_log2 >= 0 && wpos == TW - 1 && lane == 31 && base < N) a.tau_tile[tile] = tau[3];
      }
    }
    if (SH && blockIdx.x == 0 && tid == 0) *a.total = etile(a.ntiles_g - 1);   // E[Ng]
    trace(t, 6);
    if (all_phases) grid.sync();                       // (gpu-scope fence: tau and the particles of this step are in this GPU's L2,
                                                       //  which is where the peers' NVLink reads are served from)
    trace(t, 7);
    if (p2p) xbarrier(2, t);
    trace(t, 14);
    lap(3);
    w_uni = true; gather = true;
  }
  if (!(a.phases & PF_PH_FINAL)) return;

  // ------------------------------------------------------------------ materialise the carry
  const int ptf = all_phases ? T : a.t_offset;          // number of steps behind the carry
  const float* const* xpf_peers = (ptf & 1) ? a.xa_peers : a.xb_peers;
  xp_stage = xpf_peers;
  if (gather) begin_gather(T - 1);
  for (int q = next_chunk(0); q < nchunk_valid; q = next_chunk(0)) {
    const int base = pbase(q) + lane * 4;
    const int nv = max(0, min(4, N - base));
    int anc[4] = {a.goff + base, a.goff + base + 1, a.goff + base + 2, a.goff + base + 3};
    int xs_lo = -1;
    xp_stage = xpf_peers;
    if (gather) ancestors4(T - 1, q, base, nv, anc, xs_lo);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nv) {
        const int i = base + k;
        float x[DX];
        if (SH && xs_lo >= 0) {
          const float* xs = s_xst + ((size_t)(2 * warp + x_buf) * PF_FUSED_XCAP + (anc[k] - xs_lo)) * DX;
#pragma unroll
          for (int c = 0; c < DX; ++c) x[c] = xs[c];
        } else if (gather) load_prev(anc[k], x_prev, xpf_peers, x);
        else load_state_cg<DX>(x_prev, (size_t)i, x);
        store_state<DX>(a.carry_x, (size_t)i, x);
        if (a.final_particles) store_state<DX>(a.final_particles, (size_t)i, x);
        if (a.out_particles) store_state<DX>(a.out_particles + ((ptrdiff_t)T - 1) * N * DX, (size_t)i, x);
        if (a.out_ancestors) a.out_ancestors[((ptrdiff_t)T - 1) * N + i] = anc[k];
        const float w = w_uni ? invN : a.W[i];
        if (w_uni) a.W[i] = invN;
        if (a.final_weights) a.final_weights[i] = w;
      }
    }
  }
  lap(4);
  if (blockIdx.x == 0 && tid == 0) {
    if (p2p) *a.epoch = *a.epoch + 1ull;
    StepStats st = all_phases ? s_stats : a.carry_stats[a.t_offset & 1];
    st.flag = 0; st.w_uniform = w_uni ? 1 : 0;
    a.carry_stats[0] = st;
    if (!all_phases) a.carry_stats[1] = st;
  }
}
