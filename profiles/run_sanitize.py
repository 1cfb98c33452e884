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
# round 2: multinomial cell table (forced on at a small N), unscented and augmented filters, Lorenz-96 d = 6 (padded workspace)
os.environ["BF_PF_CELL_TABLE"] = "1"
c = case("bot2"); T = 4
_, ys = c.simulate(T, 1); u = c.inputs(T)
eng = ParticleFilterEngine(c.host_model(), 16384, 1, 1.0, "multinomial", "streamed", 3, 0)
out = eng.run_host(ys, inputs=u, outputs=("means", "log_likelihood", "ancestors"))
eng.close()
print("cell table", float(out["log_likelihood"].sum()))
from bayesianfiltering_b200 import inference as gf
from bayesianfiltering_b200.random import PRNGKey
c = case("bot2"); T = 5
_, ys = c.simulate(T, 2)
post = gf.unscented_gaussian_sum_filter(c.params_nlssm(), gf.ParamsUKF(1.0, 0.0, 0.0), ys, 6, 1, c.inputs(T))
print("ugsf", float(np.asarray(post.weights).sum()))
post, aux = gf.speedy_augmented_gaussian_sum_filter(c.params_nlssm(), ys, [3, 3, 3], PRNGKey(1), 1, (0.1, 0.1), c.inputs(T))
print("agsf", float(np.asarray(post.weights).sum()))
c = case("l96_6"); T = 3
_, ys = c.simulate(T, 2)
eng = GaussianSumFilterEngine(c.host_model(), 13, 2, 0)
out = eng.run_host(np.stack([ys, ys]), init_means=np.stack([c.gsf_means0(13, b) for b in range(2)]), outputs=("weights", "means"))
eng.close()
print("l96 d=6", float(out["weights"].sum()))
