import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine
T = 40; N = 1 << 24
c = case("bot2")
_, ys = c.simulate(T, 1000)
u = c.inputs(T)
eng = ParticleFilterEngine(c.host_model(), N, 1, 1.0, "systematic", "streamed", 1234, 0)
out = eng.run_host(ys, inputs=u, outputs=("ess", "log_likelihood"))
print("ess", out["ess"][0])
out = eng.run_host(ys[:3], inputs=u[:3], outputs=("norm_weights",))
w = out["norm_weights"][0, 2].astype(np.float64)
cdf = np.cumsum(w)
Tl = 2048
edges = np.searchsorted(cdf, (np.arange(0, N + 1, Tl)) / N * cdf[-1])
win = np.diff(edges)
print("window sizes: mean", win.mean(), "median", np.median(win), "max", win.max(), "frac > 6144:", (win > 6144).mean(), "p99", np.percentile(win, 99))
