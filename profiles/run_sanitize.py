"""Small particle-filter and Gaussian-sum runs for compute-sanitizer (racecheck / memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine, GaussianSumFilterEngine
for name, N, res, thr in (("bot2", 8192, "systematic", 1.0), ("ungm", 5000, "multinomial", 0.5), ("lg8", 4096, "systematic", 0.7)):
    c = case(name); T = 6
    _, ys = c.simulate(T, 1); u = c.inputs(T)
    for path in ("streamed", "resident") if name == "ungm" else ("streamed",):
        eng = ParticleFilterEngine(c.host_model(), N, 1, thr, res, path, 3, 0)
        out = eng.run_host(ys, inputs=u, outputs=("means", "log_likelihood", "ancestors", "particles"))
        eng.close()
        print(name, path, float(out["log_likelihood"].sum()))
c = case("ungm"); T = 8
_, ys = c.simulate(T, 2)
eng = GaussianSumFilterEngine(c.host_model(), 100, 8, 0)
out = eng.run_host(ys, inputs=c.inputs(T), init_means=np.stack([c.gsf_means0(100, b) for b in range(8)]), outputs=("weights", "means"))
eng.close()
print("gsf", float(out["weights"].sum()))
