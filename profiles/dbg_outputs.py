"""Debug: does the general kernel's result depend on which outputs are requested?"""
import sys, numpy as np
sys.path.insert(0, '.')
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine
ALL_OUT = ("weights", "particles", "propagated", "norm_weights", "ancestors", "means", "log_likelihood", "ess",
           "resampled", "final_particles", "final_weights")
c = case("bot2")
T, N = 8, 16384
_, ys = c.simulate(T, 7)
u = c.inputs(T)
x0, noise, ru = c.pf_randomness(T, N, 8, "systematic")
res = {}
for name, outs in (("all", ALL_OUT), ("few", ("means", "log_likelihood", "resampled")), ("few2", ("means", "log_likelihood", "resampled")),
                   ("anc", ("means", "log_likelihood", "ancestors")), ("prop", ("means", "log_likelihood", "propagated")),
                   ("part", ("means", "log_likelihood", "particles")), ("nw", ("means", "log_likelihood", "norm_weights")),
                   ("w", ("means", "log_likelihood", "weights")), ("fin", ("means", "log_likelihood", "final_particles"))):
    eng = ParticleFilterEngine(c.host_model(), N, 1, 1.0, "systematic", "streamed", 3, 0)
    o = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=outs)
    eng.close()
    res[name] = o
    print(name, np.array2string(o["log_likelihood"][0], precision=6), flush=True)
