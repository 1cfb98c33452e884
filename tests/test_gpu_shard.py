"""GPU tier, needs >= 2 GPUs (skipped on a 1-GPU box): one particle filter sharded over the GPUs of a box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_sharded_particle_filter_matches_the_oracle_and_the_one_gpu_filter(mode):
    """both transports of the per-step exchange: peer-memory stores + flag barriers inside the persistent kernel (default),
    and one launch per phase with NCCL collectives between them"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4 if n < 8 else 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, BF_SHARD_MODE=mode))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("equals the one-GPU run: ok") == 3, r.stdout[-3000:]       # three Philox runs (dist_shard_check.py)
