"""Timings of the widened rows through the reference-facing entry points, host arrays in and out:
unscented Gaussian-sum filter, speedy augmented Gaussian-sum filter, legacy UKF / MCF banks.  One B200."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.parity_cases import case
from bayesianfiltering_b200 import inference as gf
from bayesianfiltering_b200.random import PRNGKey
torch.zeros(1, device="cuda")

def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

up = gf.ParamsUKF(1.0, 0.0, 0.0)
for name, M, T, B in (("ungm", 100, 500, None), ("bot2", 100, 500, None), ("ungm", 100, 500, 1024), ("bot2", 100, 500, 100)):
    c = case(name)
    ys = c.simulate(T, 1)[1] if B is None else np.stack([c.simulate(T, 10 + b % 8)[1] for b in range(B)])
    p = c.params_nlssm()
    ms = timeit(lambda: gf.unscented_gaussian_sum_filter(p, up, ys, M, 1, c.inputs(T), num_trials=B))
    ms2 = timeit(lambda: gf.unscented_gaussian_sum_filter(p, up, ys, M, 1, c.inputs(T), num_trials=B, outputs=("weights", "means")))
    nb = 1 if B is None else B
    print(f"UGSF {name} M={M} T={T} trials={nb}: {ms:.2f} ms all five fields ({M*T*nb/(ms*1e-3):.3g} component-steps/s), "
          f"{ms2:.2f} ms weights+means ({M*T*nb/(ms2*1e-3):.3g})", flush=True)
for name, nc, T, B in (("ungm", (5, 3, 3), 100, None), ("bot2", (5, 5, 5), 100, None), ("ungm", (10, 5, 5), 100, 256), ("bot2", (5, 5, 5), 100, 256)):
    c = case(name)
    ys = c.simulate(T, 1)[1] if B is None else np.stack([c.simulate(T, 10 + b % 8)[1] for b in range(B)])
    p = c.params_nlssm()
    ms = timeit(lambda: gf.speedy_augmented_gaussian_sum_filter(p, ys, list(nc), PRNGKey(3), 1, (0.1, 0.1), c.inputs(T), num_trials=B))
    nb = 1 if B is None else B
    leaves = nc[0] * nc[1] * nc[2]
    print(f"AGSF {name} [M,N1,N2]={list(nc)} T={T} trials={nb}: {ms:.2f} ms  {leaves*T*nb/(ms*1e-3):.3g} leaf-steps/s", flush=True)
