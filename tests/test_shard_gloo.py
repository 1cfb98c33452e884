"""CPU tier: the N>1 path of ONE particle filter sharded over ranks (BASELINE configs[2]).

The CUDA path (engine.ShardedParticleFilterEngine, csrc/pf_fused.cuh) needs NVLink peers; what it exchanges and why a
sharded filter keeps the one-rank bits is restated in oracle/shard.py and checked here -- in process for 2, 4 and 8 ranks
and with a real exchange over a world-size-2 gloo group: every rank contributes ONE number to the scan (its subtree
total), scans the totals itself, finishes its own leaves, and the ancestors of its output slots equal the one-rank
`_resample` (gaussfiltax/utils.py:207-214) bit for bit.  The multi-GPU run of the kernels themselves is
tests/test_gpu_shard.py (skipped below two GPUs) and `bench.py --gpus N` (`sharded.parity_vs_1gpu`).
"""
import os
import socket

import numpy as np
import pytest

from oracle import scan, shard


def _weights(n, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "flat":
        w = rng.random(n, dtype=np.float32)
    elif kind == "peaked":                       # bearings-only-like: the mass sits in a few index ranges
        w = np.exp(-0.5 * ((np.arange(n) - 0.3 * n) / (0.01 * n)) ** 2).astype(np.float32) + 1e-12 * rng.random(n, dtype=np.float32)
    else:                                        # log-normal, five orders of magnitude
        w = np.exp(5.0 * rng.standard_normal(n)).astype(np.float32)
    return (w / w.sum(dtype=np.float32)).astype(np.float32)


@pytest.mark.parametrize("kind", ["flat", "peaked", "lognormal"])
@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("n", [8, 2048, 1 << 15])
def test_sharded_scan_has_the_one_rank_bits(n, G, kind):
    w = _weights(n, kind, n + G)
    assert np.array_equal(shard.sharded_cumsum(w, G), scan.cumsum_associative(w))


def test_offset_plus_local_prefix_would_not_do():
    """Why the down-sweep is seeded instead of adding the offset afterwards: fp32 addition is not associative."""
    w = _weights(4096, "flat", 3)
    ref = scan.cumsum_associative(w)
    S = 1024
    naive = np.concatenate([scan.cumsum_associative(w[r * S:(r + 1) * S]) + (ref[r * S - 1] if r else np.float32(0)) for r in range(4)])
    assert (naive != ref).sum() > 0
    # ... and why a rank's LAST leaf takes the canonical prefix of the totals, not offset + own total (G = 4, rank 3)
    T = np.asarray([shard.subtree_total(w[r * S:(r + 1) * S]) for r in range(4)], np.float32)
    excl, incl = shard.rank_offsets(T)
    assert incl[3] == np.float32(np.float32(T[0] + T[1]) + np.float32(T[2] + T[3])) == ref[-1]


def _worker(rank, world, port, n, mode, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = n // world
        w = _weights(n, "peaked", 11)            # the same filter on every rank; a rank only TOUCHES its own shard
        mine = w[rank * S:(rank + 1) * S]
        # exchange 1 (scan offsets): one float per rank
        tot = torch.tensor([shard.subtree_total(mine)], dtype=torch.float32)
        allt = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(allt, tot)
        excl, incl = shard.rank_offsets(torch.cat(allt).numpy())
        cdf_mine = shard.shard_cumsum(mine, excl[rank], incl[rank])
        # exchange 2 (resampling): on the GPUs a rank reads the peers' thresholds through NVLink; here, all_gather
        parts = [torch.zeros(S) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(cdf_mine))
        cdf = torch.cat(parts).numpy()
        total = cdf[-1]
        if mode == "multinomial":
            u = np.random.default_rng(5).random(n, dtype=np.float32)[rank * S:(rank + 1) * S]
            r = scan.multinomial_positions(total, u)
        else:
            r = scan.systematic_positions(total, np.float32(0.37), n)[rank * S:(rank + 1) * S]
        anc = np.minimum(scan.searchsorted_bisect(cdf, r), n - 1).astype(np.int32)
        q.put((rank, cdf_mine.tobytes(), anc.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["multinomial", "systematic"])
def test_two_ranks_exchange_totals_and_reproduce_the_one_rank_resampler(mode):
    import torch.multiprocessing as mp
    n, world = 1 << 13, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, mode, q)) for r in range(world)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    w = _weights(n, "peaked", 11)
    cdf = scan.cumsum_associative(w)
    u = np.random.default_rng(5).random(n, dtype=np.float32) if mode == "multinomial" else np.float32(0.37)
    anc = scan.resample_ancestors(w, u, mode)
    S = n // world
    for rank, cdf_b, anc_b in res:
        assert np.array_equal(np.frombuffer(cdf_b, np.float32), cdf[rank * S:(rank + 1) * S])
        assert np.array_equal(np.frombuffer(anc_b, np.int32), anc[rank * S:(rank + 1) * S])
    # the peaked weights send most of rank 1's slots to ancestors that live on rank 0 (the remote-gather case of DESIGN section 5)
    assert (anc[S:] < S).mean() > 0.5
