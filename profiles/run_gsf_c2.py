"""Small driver for ncu: BASELINE configs[1] -- GSF(EKF) on UNGM, M=100 x 1024 trials x T=500, five posterior fields."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200.engine import GaussianSumFilterEngine
M, B, T = 100, 1024, 500
g = case("ungm")
gy = np.stack([g.simulate(T, 50000 + b)[1] for b in range(16)])
gy = np.tile(gy, (B // 16, 1, 1)).astype(np.float32)
gu = g.inputs(T)
gm0 = np.stack([g.gsf_means0(M, 60000 + b) for b in range(B)])
eng = GaussianSumFilterEngine(g.host_model(), M, B, 0)
dev = torch.device("cuda", 0)
gy_d, gu_d, gm0_d = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev) for a in (gy, gu, gm0))
fields = ("weights", "means", "covariances", "predicted_means", "predicted_covariances")
for _ in range(3):
    out = eng.run(gy_d, inputs=gu_d, init_means=gm0_d, outputs=fields)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = eng.run(gy_d, inputs=gu_d, init_means=gm0_d, outputs=fields)
e1.record(); torch.cuda.synchronize()
print("ms per run", e0.elapsed_time(e1) / 5, "component-steps/s", M * B * T / (e0.elapsed_time(e1) / 5 * 1e-3), float(out["weights"].sum()))
