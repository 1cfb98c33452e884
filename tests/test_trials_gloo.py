"""CPU tier, world_size 2 over gloo: the N>1 path of the Monte-Carlo trial driver (trial sharding).
No kernels: the per-block function is a stand-in; what is tested is the partition, the gather and its order."""
import os
import socket

import numpy as np
import pytest

from bayesianfiltering_b200 import montecarlo as mc


def test_shard_range_is_a_balanced_contiguous_partition():
    for B in (0, 1, 2, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [mc.shard_range(B, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        mc.shard_range(4, 2, 2)


def test_run_trials_without_a_process_group_is_a_plain_call():
    out = mc.run_trials(lambda a, b: {"x": np.arange(a, b)}, 5)
    assert np.array_equal(out["x"], np.arange(5))


def test_rmse_and_bootstrap_ci():
    est = np.zeros((3, 4, 2)); tru = np.ones((3, 4, 2))
    np.testing.assert_allclose(mc.rmse(est, tru), np.sqrt(2.0))          # utils.py:179-187 per trial
    m, (lo, hi) = mc.bootstrap_ci([1.0, 2.0, 3.0, np.nan])
    assert m == 2.0 and lo <= 2.0 <= hi


def _worker(rank, world, port, B, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def block(start, stop):
            calls.append((start, stop))
            t = np.arange(start, stop, dtype=np.float64)
            return {"marginal_log_likelihood": -t, "means": np.stack([t, 2 * t], axis=1)[:, None, :].repeat(3, axis=1)}

        out = mc.run_trials(block, B)
        q.put((rank, calls, out["marginal_log_likelihood"].tolist(), out["means"].shape, out["means"][:, 0, 1].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 2, 1])
def test_trials_shard_over_two_ranks_and_gather_in_trial_order(B):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    starts = [mc.shard_range(B, 2, r) for r in range(2)]
    for rank, calls, ll, shape, m1 in res:
        assert calls == ([starts[rank]] if starts[rank][1] > starts[rank][0] else [])     # each rank ran only its own block
        assert ll == [-float(i) for i in range(B)]                                        # every rank sees all trials, in order
        assert tuple(shape) == (B, 3, 2) and m1 == [2.0 * i for i in range(B)]
