"""Phase times of the fused particle filter (BOT, N=2^24, systematic, Philox): python profiles/run_pf_phases.py [T] [log2N] [resampler] [bot|lg8]
BFILT_LIB selects the library under test."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayesianfiltering_b200 import workloads
from bayesianfiltering_b200.engine import ParticleFilterEngine
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
RES = sys.argv[3] if len(sys.argv) > 3 else "systematic"
WL = sys.argv[4] if len(sys.argv) > 4 else "bot"
wl = getattr(workloads, WL)()
_, ys = wl.sample(T, 1000)
u = wl.inputs(T)
dev = torch.device("cuda", 0)
eng = ParticleFilterEngine(wl.host_model(), N, 1, 1.0, RES, "streamed", 1234, 0) if WL == "bot" else ParticleFilterEngine(wl.host_model(), N, 1, 0.5, RES, "streamed", 1234, 0)
ys_d = torch.from_numpy(ys).to(dev)
u_d = torch.from_numpy(u).to(dev) if u is not None else None
for _ in range(2):
    out = eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
torch.cuda.synchronize()
eng.phase_times(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = eng.run(ys_d, inputs=u_d, outputs=("means", "log_likelihood"))
e1.record(); torch.cuda.synchronize()
ph = {k: round(v * 1e3 / (3 * T), 1) for k, v in eng.phase_times().items()}
print(os.environ.get("BFILT_LIB", "default"), WL, RES, "cells=" + os.environ.get("BF_PF_CELL_TABLE", "auto"), "us/step", round(e0.elapsed_time(e1) * 1e3 / (3 * T), 1), ph, "loglik", float(out["log_likelihood"].double().sum()))
