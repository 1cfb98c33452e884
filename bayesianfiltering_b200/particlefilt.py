"""Legacy bootstrap particle filter with the call surface of gaussfiltax/particlefilt.py:11-52, on libbfilt.so.

`BootstrapPF(ssm, N).run(ys, m0, P0)` -> particles [T+1, N, dx] (initial particles at index T, particlefilt.py:29-30).
The reference resamples at every step with multinomial COUNTS (:45-50), i.e. multinomial resampling whose
ancestors come out sorted; weights are therefore never carried.  Here: ess_threshold > 1 (always resample) and
multinomial positions from sorted uniforms.  The reference draws from NumPy's global generator, so only the
distribution can match -- unless `init_particles`, `noise`, `resample_uniforms` are supplied.
"""
import time

import numpy as np
from numpy import random

from .engine import ParticleFilterEngine


class BootstrapPF:

    def __init__(self, ssm, N):
        self.ssm = ssm
        self.f, self.g, self.Q, self.R, self.dx, self.dy = ssm.f, ssm.g, ssm.Q, ssm.R, ssm.dx, ssm.dy
        self.N = N
        self.time = 0.0

    def __str__(self):
        return 'BPF'

    def run(self, ys, m0, P0, verbose=False, init_particles=None, noise=None, resample_uniforms=None, path="auto"):
        tin = time.time()
        T = np.shape(ys)[0]
        if init_particles is None:           # particlefilt.py:30
            init_particles = random.multivariate_normal(np.asarray(m0, np.float64).reshape(-1), np.asarray(P0, np.float64), self.N)
        if resample_uniforms is None:        # multinomial counts == ancestors of sorted uniforms
            resample_uniforms = 1.0 - np.sort(random.random((T, self.N)), axis=1)      # positions total*(1-u) ascending
        _, hm = self.ssm._host_model(m0, P0)
        eng = ParticleFilterEngine(hm, self.N, 1, 2.0, "multinomial", path, int(random.randint(0, 2 ** 31 - 1)))
        out = eng.run_host(np.asarray(ys, np.float32).reshape(T, self.dy), init_particles=np.asarray(init_particles, np.float32),
                           noise=noise, resample_u=np.asarray(resample_uniforms, np.float32), outputs=("particles",))
        eng.close()
        particles = np.zeros((T + 1, self.N, self.dx))
        particles[:T] = out["particles"][0]
        particles[T] = init_particles
        self.time = time.time() - tin
        return particles
