"""BASELINE configs[4]: particle-filter sweep N = 2^10 .. 2^28 on a linear-Gaussian SSM (d = 8), validated against the
exact Kalman-filter log-likelihood (float64, oracle/gsf.py kalman_filter_exact).
usage: python profiles/run_c5_sweep.py [max_log2N] [T]   ->  one JSON line per N on stdout"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import gsf as ogsf
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine

max_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 26
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
c = case("lg8")
_, ys = c.simulate(T, 9)
_, _, ll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
exact = float(ll.sum())
dev = torch.device("cuda", 0)
ys_d = torch.from_numpy(ys).to(dev)
for log2n in range(10, max_log2 + 1, 2):
    N = 1 << log2n
    vals, ms = [], []
    for seed in range(3):
        eng = ParticleFilterEngine(c.host_model(), N, 1, 0.5, "systematic", "auto", 100 + seed, 0)
        out = eng.run(ys_d, outputs=("log_likelihood",))                     # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.run(ys_d, outputs=("log_likelihood",))
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        vals.append(float(out["log_likelihood"].double().sum()))
        eng.close()
    err = np.abs(np.array(vals) - exact)
    print(json.dumps({"log2N": log2n, "T": T, "logZ_kalman": exact, "logZ_pf": vals, "abs_err_mean": float(err.mean()),
                      "ms": float(np.median(ms)), "particle_steps_per_s": N * T / (np.median(ms) * 1e-3)}), flush=True)
