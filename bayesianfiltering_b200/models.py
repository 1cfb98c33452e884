"""Model and parameter containers, same names and fields as gaussfiltax/models.py.

`ParamsNLSSM` (models.py:26-51) and `ParamsBPF` (models.py:55-84) are NamedTuples of arrays plus the
model callables; here the callables must come from the compiled registry (`bayesianfiltering_b200.zoo`).
`NonlinearSSM.sample` (models.py:240-289) is the data generator; it runs on the GPU (`bf_ssm_sample`).
"""
import ctypes as C
from typing import Any, Callable, NamedTuple

import numpy as np

from . import _lib, zoo
from .random import key_to_seed
from .ssm import SSM


class ParamsNLSSM(NamedTuple):
    """Parameters for a NLGSSM model (gaussfiltax/models.py:26-51)."""
    initial_mean: Any
    initial_covariance: Any
    dynamics_function: Callable
    dynamics_noise_bias: Any
    dynamics_noise_covariance: Any
    emission_function: Callable
    emission_noise_bias: Any
    emission_noise_covariance: Any


class ParamsBPF(NamedTuple):
    """Parameters for a BPF model (gaussfiltax/models.py:55-84)."""
    initial_mean: Any
    initial_covariance: Any
    dynamics_function: Callable
    dynamics_noise_bias: Any
    dynamics_noise_covariance: Any
    emission_function: Callable
    emission_noise_bias: Any
    emission_noise_covariance: Any
    emission_distribution_log_prob: Callable

    def sample_dynamics_distribution(self, key, x, u):
        """q ~ MVN(q0, Q); return f(x, q, u)   (models.py:82-84), evaluated on the GPU."""
        rng = np.random.default_rng(key_to_seed(key))
        Q = np.asarray(self.dynamics_noise_covariance, dtype=np.float32)
        L = np.linalg.cholesky(Q)
        q = np.asarray(self.dynamics_noise_bias, dtype=np.float32) + L @ rng.standard_normal(Q.shape[0]).astype(np.float32)
        return self.dynamics_function(x, q, u)


class NonlinearSSM(SSM):
    """General nonlinear state-space model z_t = f(z_{t-1}, q_t, u_t), y_t = h(z_t, r_t, u_t)
    (gaussfiltax/models.py:160-289)."""

    def __init__(self, state_dim: int, state_noise_dim: int, emission_dim: int, emission_noise_dim: int, input_dim: int = 0):
        self.state_dim = state_dim
        self.state_noise_dim = state_noise_dim
        self.emission_dim = emission_dim
        self.emission_noise_dim = emission_noise_dim
        self.input_dim = 0          # sic: models.py:200

    @property
    def emission_shape(self):
        return (self.emission_dim,)

    @property
    def inputs_shape(self):
        return (self.input_dim,) if self.input_dim > 0 else None

    def sample(self, params, key, num_timesteps, inputs=None, num_trials=None, z0=None, q_noise=None, r_noise=None):
        """Sample states z_{1:T} and emissions y_{1:T} (models.py:240-289).

        Returns NumPy arrays (states [T,dx], emissions [T,dy]); with `num_trials=B` a leading B axis.
        The draws come from Philox(key) unless already-coloured `z0` [B,dx], `q_noise` [B,T,dq] (row 0
        unused) and `r_noise` [B,T,dr] are supplied.
        """
        import torch
        lib = _lib.load()
        m, hm = zoo.host_model_from_params(params)
        if (m.dx, m.dq, m.dy, m.dr) != (self.state_dim, self.state_noise_dim, self.emission_dim, self.emission_noise_dim):
            raise ValueError("NonlinearSSM dims do not match the registry model in params")
        B = 1 if num_trials is None else int(num_trials)
        T = int(num_timesteps)
        dev = torch.device("cuda", torch.cuda.current_device())
        up = lambda a, shape: None if a is None else torch.as_tensor(
            np.ascontiguousarray(np.asarray(a, dtype=np.float32)).reshape(shape)).to(dev)
        u_d = up(inputs, (T,))
        z0_d, q_d, r_d = up(z0, (B, m.dx)), up(q_noise, (B, T, m.dq)), up(r_noise, (B, T, m.dr))
        states = torch.empty((B, T, m.dx), dtype=torch.float32, device=dev)
        emissions = torch.empty((B, T, m.dy), dtype=torch.float32, device=dev)
        p = lambda t: None if t is None else t.data_ptr()
        _lib.check(lib.bf_ssm_sample(C.byref(hm.desc), B, T, p(u_d), key_to_seed(key), p(z0_d), p(q_d), p(r_d),
                                     states.data_ptr(), emissions.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        s, e = states.cpu().numpy(), emissions.cpu().numpy()
        return (s[0], e[0]) if num_trials is None else (s, e)
