"""The named workloads of BASELINE.json `configs`, as registry params + synthetic data.  Used by bench.py and profiles/;
nothing here touches oracle/ or tests/: emissions come from the package's own generator (`NonlinearSSM.sample`,
models.py:240-289 -> bf_ssm_sample on the GPU).

  bot()   configs[2]  manoeuvring bearings-only tracking, dx=4 dq=2 dy=2 (BOT_Experiment_script.py:19-53)
  ungm()  configs[0/1] univariate nonstationary growth model with u_t = 8 cos(1.2 t) (Experiment_TSP_2023.ipynb cell 2)
  l96()   configs[3]  Lorenz-96 d=40, every second state observed through c1 x + c2 x^2
  lg8()   configs[4]  linear-Gaussian d=8
"""
import numpy as np

from . import zoo
from .models import NonlinearSSM, ParamsBPF, ParamsNLSSM


class Workload:
    def __init__(self, name, model, m0, P0, Q, R, inputs_fn=None):
        f = lambda a: np.asarray(a, dtype=np.float32)
        self.name, self.model = name, model
        self.m0, self.P0, self.Q, self.R = f(m0), f(P0), f(Q), f(R)
        self.q0, self.r0 = np.zeros(model.dq, np.float32), np.zeros(model.dr, np.float32)
        self.inputs_fn = inputs_fn

    def inputs(self, T):
        return None if self.inputs_fn is None else np.asarray(self.inputs_fn(T), dtype=np.float32)

    def params_nlssm(self):
        z = self.model
        return ParamsNLSSM(self.m0, self.P0, z.dynamics_function, self.q0, self.Q, z.emission_function, self.r0, self.R)

    def params_bpf(self):
        z = self.model
        return ParamsBPF(self.m0, self.P0, z.dynamics_function, self.q0, self.Q, z.emission_function, self.r0, self.R,
                         z.emission_log_prob)

    def host_model(self):
        return zoo.host_model_from_params(self.params_nlssm())[1]

    def sample(self, T, seed, num_trials=None):
        """states, emissions of T steps ([B,T,.] with num_trials) from the GPU data generator."""
        z = self.model
        return NonlinearSSM(z.dx, z.dq, z.dy, z.dr).sample(self.params_nlssm(), seed, T, inputs=self.inputs(T), num_trials=num_trials)

    def init_means(self, M, B, seed):
        """[B,M,dx] draws of N(m0, P0): the initial component means of a Gaussian-sum filter (inference.py:367)."""
        rng = np.random.default_rng(seed)
        L = np.linalg.cholesky(self.P0.astype(np.float64))
        return (self.m0 + rng.standard_normal((B, M, self.model.dx)) @ L.T).astype(np.float32)


def bot(emission_dim=2):
    maneuver = lambda T: np.array([1] * int(2 * T / 5) + [0] * int(T / 5) + [2] * (T - int(2 * T / 5) - int(T / 5)))
    return Workload("bot", zoo.bot(emission_dim=emission_dim), [-0.05, 0.001, 0.7, -0.05], np.diag([0.1, 0.005, 0.1, 0.01]),
                    1e-5 * np.eye(2), 25e-6 * np.eye(emission_dim), maneuver)


def ungm():
    return Workload("ungm", zoo.ungm(), [0.0], [[1.0]], [[1.0]], [[1.0]], lambda T: 8.0 * np.cos(1.2 * np.arange(T)))


def l96(d=40):
    return Workload("l96", zoo.lorenz96(d), 8.0 + 0.5 * np.sin(np.arange(d)), 0.5 * np.eye(d), 0.1 * np.eye(d), 0.5 * np.eye(d // 2))


def lg8():
    A, H = 0.8 * np.eye(8) + 0.05 * np.eye(8, k=1), 0.5 * np.eye(8)
    w = Workload("lg8", zoo.linear_gaussian(A, H), np.zeros(8), np.eye(8), np.eye(8), 0.1 * np.eye(8))
    w.A, w.H = A, H                                   # (for the exact Kalman yardstick of the sweep)
    return w
