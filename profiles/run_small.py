"""The reference's own problem sizes through the reference-facing entry points (host arrays in, host arrays out):
config 1 (UNGM, N=1000, T=100), the BOT experiment (N=10^4, T=500, BOT_Experiment_script.py) and a batch of trials."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200 import inference as gf
from bayesianfiltering_b200.random import PRNGKey
torch.zeros(1, device="cuda")

def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

for name, N, T, thr, B in (("ungm", 1000, 100, 0.5, None), ("bot2", 10000, 500, 1.0, None), ("bot2", 50000, 500, 1.0, None),
                           ("bot2", 10000, 500, 1.0, 100), ("ungm", 1000, 100, 0.5, 1024)):
    c = case(name)
    ys = c.simulate(T, 1)[1] if B is None else np.stack([c.simulate(T, 10 + b % 8)[1] for b in range(B)])
    u = c.inputs(T)
    p = c.params_bpf()
    for outs in (("weights", "particles"), ("means", "log_likelihood")):
        ms = timeit(lambda: gf.bootstrap_particle_filter(p, ys, N, PRNGKey(1), u, thr, outputs=outs, num_trials=B))
        nb = 1 if B is None else B
        print(f"BPF {name} N={N} T={T} trials={nb} outputs={outs[0]}..: {ms:.2f} ms  {N*T*nb/(ms*1e-3):.3g} particle-steps/s", flush=True)
for name, M, T, B in (("ungm", 100, 500, None), ("bot2", 100, 500, None), ("bot2", 100, 500, 100)):
    c = case(name)
    ys = c.simulate(T, 1)[1] if B is None else np.stack([c.simulate(T, 10 + b % 8)[1] for b in range(B)])
    p = c.params_nlssm()
    ms = timeit(lambda: gf.gaussian_sum_filter(p, ys, M, 1, c.inputs(T), num_trials=B))
    nb = 1 if B is None else B
    print(f"GSF {name} M={M} T={T} trials={nb}: {ms:.2f} ms  {M*T*nb/(ms*1e-3):.3g} component-steps/s", flush=True)
