"""Canonical prefix sum and bisection search used by the resampler.  TEST INFRASTRUCTURE ONLY.

Restates the third-party pieces behind `gaussfiltax/utils.py:207-214` (`_resample`), which is
`jax.random.choice(key, arange(N), (N,), p=w)`:

    p_cuml = jnp.cumsum(p)                     # lax.associative_scan(add) on CPU/GPU
    r      = p_cuml[-1] * (1 - uniform(key))   # fp32
    ind    = jnp.searchsorted(p_cuml, r)       # side='left', method='scan' (fixed bisection)

jax / jaxlib are NOT vendored by the reference (unpinned, `setup.py:7-10`); the published
algorithms are restated here and are the definition our CUDA kernels are bit-exact against.
"""
import numpy as np


def cumsum_associative(x):
    """Inclusive prefix sum in the summation order of `lax.associative_scan(add, x)`.

    Published algorithm (jax/_src/lax/control_flow/loops.py `associative_scan._scan`):
      reduced = x[0:-1:2] + x[1::2]            # adjacent pairs
      odd     = _scan(reduced)                 # inclusive prefixes at odd positions
      even[k] = odd[k-1] + x[2k]   (k >= 1),  even[0] = x[0]
      result  = interleave(even, odd)
    Floating-point addition is not associative, so this fixes every bit of the result.
    """
    x = np.asarray(x)
    n = x.shape[0]
    if n < 2:
        return x.copy()
    reduced = x[0:n - 1:2] + x[1::2]
    odd = cumsum_associative(reduced)
    if n % 2 == 0:
        even_tail = odd[:-1] + x[2::2]
    else:
        even_tail = odd + x[2::2]
    out = np.empty_like(x)
    out[0] = x[0]
    out[2::2] = even_tail
    out[1::2] = odd
    return out


def cumsum_closed_form(x):
    """Same numbers as `cumsum_associative`, written the way the CUDA kernel computes them.

    prefix[i] = left-to-right sum, from the most- to the least-significant set bit of n=i+1,
    of the balanced pairwise-tree sums of the aligned power-of-two blocks tiling [0, n).
    Equivalently an exclusive Blelloch down-sweep E(right) = E(parent) + T(left) and
    prefix[i] = E(leaf i+1).  O(N log N) python-free, used by tests on small N only.
    """
    x = np.asarray(x)
    n = x.shape[0]
    dt = x.dtype
    # tree sums per level: T[l][k] = sum of block [k*2^l, (k+1)*2^l)
    levels = [x.copy()]
    while levels[-1].shape[0] >= 2:
        cur = levels[-1]
        m = cur.shape[0] // 2
        levels.append((cur[0:2 * m:2] + cur[1:2 * m:2]).astype(dt))
    out = np.zeros(n, dtype=dt)
    for i in range(n):
        cnt = i + 1
        acc = None
        for l in range(len(levels) - 1, -1, -1):
            if (cnt >> l) & 1:
                k = (cnt >> l) - 1
                t = levels[l][k]
                acc = t if acc is None else dt.type(acc + t)
        out[i] = acc
    return out


def searchsorted_bisect(cdf, r):
    """`jnp.searchsorted(cdf, r, side='left')`, method='scan': a FIXED bisection.

    Published algorithm (jax/_src/numpy/lax_numpy.py `_searchsorted_via_scan`):
      low, high = 0, n ; repeat ceil(log2(n+1)) times:
         mid = (low + high) // 2 ; go_left = (r <= cdf[mid])
         low, high = (low, mid) if go_left else (mid, high)
      return high
    On a non-decreasing `cdf` this equals numpy's searchsorted(side='left') except that it can
    never return an index past the first probe sequence; the tree-ordered fp32 cdf is NOT
    guaranteed monotone in its last bit, so the probe sequence itself is part of the contract.
    """
    cdf = np.asarray(cdf)
    r = np.asarray(r)
    n = cdf.shape[0]
    n_levels = int(np.ceil(np.log2(n + 1)))
    low = np.zeros(r.shape, dtype=np.int64)
    high = np.full(r.shape, n, dtype=np.int64)
    for _ in range(n_levels):
        mid = (low + high) // 2
        # mid == n can only happen when low == high == n, which the loop never reaches
        go_left = r <= cdf[np.minimum(mid, n - 1)]
        high = np.where(go_left, mid, high)
        low = np.where(go_left, low, mid)
    return high.astype(np.int32)


def effective_thresholds_pow2(cdf):
    """Monotone thresholds tau with  searchsorted_bisect(cdf, r) == np.searchsorted(tau, r, 'left')  for EVERY r,
    for len(cdf) = n = 2^k -- including cdfs that are not monotone (the tree-ordered fp32 cumsum is not).

    Why: the fixed bisection F(r) is monotone in r whatever the array holds (two queries share their path until
    the first node where the smaller goes left and the larger goes right), so it is a lower bound on SOME
    monotone array.  Unrolling `F(r) <= j` down the bisection tree: at a node with probe value v = cdf[mid],
        j <  mid :  F <= j  needs "go left"  (r <= v) AND the left subtree's condition   ->  tau_j = min(v, tau_j(left))
        j >= mid :  F <= j  holds if "go left" OR the right subtree's condition          ->  tau_j = max(v, tau_j(right))
    with tau_j = cdf[j] at the bottom.  The CUDA kernel stores tau instead of the cdf (same bytes) and can then
    use ANY lower-bound search (windowed, galloping) and still return the reference's ancestors bit for bit.
    O(n log n), test use only."""
    cdf = np.asarray(cdf)
    n = cdf.shape[0]
    assert n >= 1 and (n & (n - 1)) == 0, "power-of-two length only"
    j = np.arange(n)
    tau = cdf.copy()
    s = 1
    while s < n:                                  # deepest ancestor first, the root last (outermost)
        mid = (j // (2 * s)) * (2 * s) + s
        v = cdf[mid]
        tau = np.where(j < mid, np.minimum(v, tau), np.maximum(v, tau))
        s *= 2
    return tau


def multinomial_positions(total, u):
    """r_i = p_cuml[-1] * (1 - u_i), u_i in [0,1)   (jax.random.choice, replace=True, p given)."""
    u = np.asarray(u)
    one = u.dtype.type(1)
    return (total * (one - u)).astype(u.dtype)


def systematic_positions(total, u0, n):
    """r_i = total * ((i + u0) / n): the systematic grid `north_star` names (not in the reference).

    fp32 op order is part of the contract: float(i) + u0, then / float(n), then * total.
    """
    dt = np.asarray(total).dtype
    i = np.arange(n).astype(dt)
    return (dt.type(total) * ((i + dt.type(u0)) / dt.type(n))).astype(dt)


def resample_ancestors(w, u, mode="multinomial"):
    """Ancestor indices of `utils._resample` (`utils.py:207-214`) for supplied uniforms.

    mode 'multinomial': u has shape [N] (one uniform per output slot, as jr.choice draws);
    mode 'systematic' : u is a scalar.
    Returns int32 [N].  Index N can be produced only if r > cdf[N-1], which
    0 < r <= cdf[-1] rules out; we clamp to N-1 like `jnp.take`'s default clip mode.
    """
    w = np.asarray(w)
    n = w.shape[0]
    cdf = cumsum_associative(w)
    if mode == "multinomial":
        r = multinomial_positions(cdf[-1], np.asarray(u, dtype=w.dtype))
    elif mode == "systematic":
        r = systematic_positions(cdf[-1], u, n)
    else:
        raise ValueError(mode)
    idx = searchsorted_bisect(cdf, r)
    return np.minimum(idx, n - 1).astype(np.int32)


def cell_table_ancestors(w, u):
    """The multinomial ancestors again, computed the way the large-N kernel does from N = 2^18 on (csrc/pf_fused.cuh,
    `begin_gather` + the `use_cells` branch of the ancestor search) -- a runnable statement of why that shortcut is exact.

      tau     = effective_thresholds_pow2(cumsum_associative(w))            monotone, same lower bounds as the bisection
      L[c]    = lower_bound(tau, total * (c / N)),  c = 0 .. N - 1;  L[N] = N - 1      the CELL TABLE (sorted queries)
      r       = total * (1 - u),  cell c = floor((1 - u) * N)                           exact: N is a power of two
      answer  = L[c] + lower_bound(tau[L[c] : L[c + 1]], r)        (L[c] itself when (1 - u) * N is an integer: r IS the edge)

    c / N <= 1 - u < (c + 1) / N and rounding a product by `total` is monotone, so edge_c <= r <= edge_{c+1} and, tau being
    sorted, L[c] <= lower_bound(tau, r) <= L[c + 1].  tests/test_oracle.py holds this to `resample_ancestors`."""
    w = np.asarray(w)
    n = w.shape[0]
    dt = w.dtype
    cdf = cumsum_associative(w)
    tau = effective_thresholds_pow2(cdf)
    total = cdf[-1]
    edges = systematic_positions(total, dt.type(0), n)
    L = np.empty(n + 1, dtype=np.int64)
    L[:n] = np.minimum(np.searchsorted(tau, edges, side="left"), n - 1)
    L[n] = n - 1
    one_minus_u = (dt.type(1) - np.asarray(u, dtype=dt)).astype(dt)
    wq = (one_minus_u * dt.type(n)).astype(dt)
    c = np.clip(wq.astype(np.int64), 0, n - 1)
    r = multinomial_positions(total, np.asarray(u, dtype=dt))
    out = np.empty(r.shape[0], dtype=np.int32)
    for i in range(r.shape[0]):
        lo = L[c[i]]
        if wq[i] == dt.type(c[i]):
            out[i] = lo
            continue
        hi = L[c[i] + 1]
        out[i] = min(lo + int(np.searchsorted(tau[lo:hi], r[i], side="left")), n - 1)
    return out
