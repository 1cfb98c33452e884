import sys, time; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine
c = case("bot2"); T = 1000; N = 1 << 24
_, ys = c.simulate(T, 1000); u = c.inputs(T)
torch.cuda.init(); torch.zeros(1, device="cuda")
for rep in range(3):
    t0 = time.perf_counter(); hm = c.host_model(); t1 = time.perf_counter()
    eng = ParticleFilterEngine(hm, N, 1, 1.0, "systematic", "streamed", 7, 0); torch.cuda.synchronize(); t2 = time.perf_counter()
    out = eng.run_host(ys, inputs=u, outputs=("means", "log_likelihood")); t3 = time.perf_counter()
    eng.close(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"host_model {1e3*(t1-t0):.1f} ms  create {1e3*(t2-t1):.1f} ms  run_host {1e3*(t3-t2):.1f} ms  close {1e3*(t4-t3):.1f} ms")
