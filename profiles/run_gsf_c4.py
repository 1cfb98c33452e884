"""BASELINE configs[3], one GPU's share: GSF(EKF) on Lorenz-96 d=40 (dy=20), M=256 components x 512 trials (4096 / 8 GPUs),
outputs = weights + mixture mean.  usage: python profiles/run_gsf_c4.py [T] [B] [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200.engine import GaussianSumFilterEngine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
M = int(sys.argv[3]) if len(sys.argv) > 3 else 256
c = case("l96")
ys = np.stack([c.simulate(T, 700 + b)[1] for b in range(8)])
ys = np.tile(ys, (B // 8, 1, 1)).astype(np.float32)
rng = np.random.default_rng(0)
m0 = (c.m0 + rng.standard_normal((B, M, 40)) @ np.linalg.cholesky(c.P0.astype(np.float64)).T).astype(np.float32)
dev = torch.device("cuda", 0)
eng = GaussianSumFilterEngine(c.host_model(), M, B, 0)
ys_d, m0_d = torch.from_numpy(ys).to(dev), torch.from_numpy(m0).to(dev)
outs = ("weights", "mixture_mean", "log_likelihood")
out = eng.run(ys_d, init_means=m0_d, outputs=outs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = eng.run(ys_d, init_means=m0_d, outputs=outs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
cs = M * B * T
print(f"ms {ms:.2f}  component-steps/s {cs / (ms * 1e-3):.4g}  carry GB/s {cs * 8 * (40 + 1600) / (ms * 1e-3) / 1e9:.1f}  "
      f"loglik {float(out['log_likelihood'].double().sum()):.3f}  wsum {float(out['weights'].sum()) / (B * T):.6f}")
