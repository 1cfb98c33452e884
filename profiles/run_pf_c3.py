"""Small driver for ncu: the bench workload (BOT, N=2^24, systematic, Philox) for a few time steps.
usage: python profiles/run_pf_c3.py [T] [log2N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200.engine import ParticleFilterEngine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
c = case("bot2")
_, ys = c.simulate(T, 1000)
u = c.inputs(T)
eng = ParticleFilterEngine(c.host_model(), N, 1, 1.0, "systematic", "streamed", 1234, 0)
dev = torch.device("cuda", 0)
ys_d, u_d = torch.from_numpy(ys).to(dev), torch.from_numpy(u).to(dev)
for _ in range(2):
    out = eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
torch.cuda.synchronize()
print("loglik", float(out["log_likelihood"].double().sum()), eng.phase_times())
