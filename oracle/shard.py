"""One particle filter sharded over G ranks: what the ranks exchange and why the result has the one-rank bits.
TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

The reference has no multi-device path; `north_star` asks for one ("NCCL over NVLink carrying only the global weight-sum
and scan-offset exchange that resampling needs") with ancestors bit-exact against the one-device `_resample`
(`gaussfiltax/utils.py:207-214` = cumsum in `lax.associative_scan` order + fixed bisection, restated in oracle/scan.py).
This file restates the decomposition the CUDA path uses (csrc/pf_fused.cuh, phases C1/C2 of a sharded filter) at the
granularity of a rank:

  * rank r owns the leaves [r S, (r+1) S), S = N / G a power of two: an ALIGNED subtree of the scan tree;
  * the ranks exchange one number each, the balanced pairwise sum T_r of their subtree (`subtree_total`);
  * every rank scans the G totals itself in the canonical order (`rank_offsets`) and finishes its own leaves with a
    down-sweep seeded by its exclusive offset (`shard_cumsum`): E(right child) = E(parent) + T(left child).

`shard_cumsum` over all ranks, concatenated, equals `scan.cumsum_associative` of the whole array bit for bit
(tests/test_shard_gloo.py: in process for G = 2, 4, 8 and over a world-size-2 gloo group).
"""
import numpy as np

from . import scan


def subtree_total(x):
    """Balanced pairwise sum of a power-of-two block -- the root of the scan tree above it."""
    x = np.asarray(x)
    assert x.shape[0] >= 1 and (x.shape[0] & (x.shape[0] - 1)) == 0, "power-of-two length only"
    while x.shape[0] > 1:
        x = x[0::2] + x[1::2]
    return x[0]


def rank_offsets(totals):
    """From the G exchanged subtree totals: (exclusive offset of every rank, inclusive prefix at every rank's last leaf),
    both in the canonical order.  The exclusive offset of rank r is the canonical inclusive prefix of the totals at r - 1
    (the bits of the count r S above log2 S are the bits of r); rank 0 has none (`None`: nothing is added)."""
    totals = np.asarray(totals)
    incl = scan.cumsum_associative(totals)
    excl = [None] + [incl[r - 1] for r in range(1, totals.shape[0])]
    return excl, incl


def shard_cumsum(local, offset, end_prefix):
    """Inclusive canonical prefix sums of one rank's leaves, given its exclusive `offset` (None on rank 0) and the canonical
    prefix at its last leaf `end_prefix` (the count (r+1) S carries into the rank-level bits, so that one element is not
    offset + T_r in general: for G = 4, r = 3 it is (T0 + T1) + (T2 + T3), not ((T0 + T1) + T2) + T3).

    Down-sweep over the rank's own tree: the exclusive prefix E of a node is handed to its left child unchanged and to its
    right child as E + T(left); the inclusive prefix of leaf i is E(leaf i + 1)."""
    local = np.asarray(local)
    S = local.shape[0]
    dt = local.dtype
    levels = [local]
    while levels[-1].shape[0] > 1:
        cur = levels[-1]
        levels.append((cur[0::2] + cur[1::2]).astype(dt))
    # E per level, top-down.  have[k] says whether node k has anything to its left at all (rank 0's leftmost spine has not).
    E = np.zeros(1, dtype=dt) if offset is None else np.asarray([offset], dtype=dt)
    have = np.asarray([offset is not None])
    for l in range(len(levels) - 2, -1, -1):
        T = levels[l]
        left_T = T[0::2]
        nE = np.empty(T.shape[0], dtype=dt)
        nhave = np.empty(T.shape[0], dtype=bool)
        nE[0::2] = E
        nhave[0::2] = have
        nE[1::2] = np.where(have, (E + left_T).astype(dt), left_T)
        nhave[1::2] = True
        E, have = nE, nhave
    # E is now the exclusive prefix of every leaf; inclusive prefix of leaf i = exclusive prefix of leaf i + 1
    out = np.empty(S, dtype=dt)
    out[:-1] = E[1:]
    out[-1] = end_prefix
    return out


def sharded_cumsum(w, G):
    """All G ranks in one process: the concatenation of their `shard_cumsum`s."""
    w = np.asarray(w)
    S = w.shape[0] // G
    totals = np.asarray([subtree_total(w[r * S:(r + 1) * S]) for r in range(G)], dtype=w.dtype)
    excl, incl = rank_offsets(totals)
    return np.concatenate([shard_cumsum(w[r * S:(r + 1) * S], excl[r], incl[r]) for r in range(G)])
