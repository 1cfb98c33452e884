"""TEST / BASELINE INFRASTRUCTURE -- not part of the product.  The CPU arm of bench.py (`cpu_baseline`, `--impl reference`)
times the oracle's bootstrap particle filter (oracle/pf.py, the restatement of inference.py:1302-1380) on the bench workload:
the manoeuvring bearings-only model with the constants of BOT_Experiment_script.py:16-53.  This module builds that workload
for the oracle (float64 simulation of the data, pre-drawn noise and uniforms) so that bench.py needs nothing from tests/."""
import time

import numpy as np

from . import models as om
from . import pf as opf

M0 = np.array([-0.05, 0.001, 0.7, -0.05], np.float32)                 # BOT_Experiment_script.py:36-37
P0 = np.diag([0.1, 0.005, 0.1, 0.01]).astype(np.float32)
Q = (1e-5 * np.eye(2)).astype(np.float32)                             # :30-33
R = (25e-6 * np.eye(2)).astype(np.float32)


def maneuver(T):                                                      # :42-47, the input that selects straight / left / right
    return np.array([1] * int(2 * T / 5) + [0] * int(T / 5) + [2] * (T - int(2 * T / 5) - int(T / 5)), dtype=np.float32)


def run_bot(seed, N, T):
    """One oracle filter: N particles, T steps, resampling at every step (ess_threshold = 1.0, :151), systematic positions.
    Returns (seconds inside the filter, marginal log-likelihood)."""
    m = om.BOT(dy=2)
    rng = np.random.default_rng(1000 + seed)
    LQ, LR, LP = (np.linalg.cholesky(a.astype(np.float64)) for a in (Q, R, P0))
    u = maneuver(T)
    x = (M0.astype(np.float64) + LP @ rng.standard_normal(4))[None]
    ys = []
    for t in range(T):
        if t > 0:
            x = m.f(x, (LQ @ rng.standard_normal(2))[None], float(u[t]))
        ys.append(m.h(x, (LR @ rng.standard_normal(2))[None], float(u[t]))[0])
    ys = np.stack(ys).astype(np.float32)
    rng = np.random.default_rng(2000 + seed)
    x0 = (M0 + rng.standard_normal((N, 4)) @ LP.T).astype(np.float32)
    noise = (rng.standard_normal((T, N, 2)) @ LQ.T).astype(np.float32)
    ru = rng.random((T,), dtype=np.float32)
    t0 = time.perf_counter()
    o = opf.bootstrap_particle_filter(m, ys, u, x0, noise, ru, np.zeros(2, np.float32), R, 1.0, "systematic", keep_trajectory=False)
    return time.perf_counter() - t0, float(o["marginal_loglik"])
