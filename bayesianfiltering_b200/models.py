"""Model and parameter containers, same names and fields as gaussfiltax/models.py.

`ParamsNLSSM` (models.py:26-51) and `ParamsBPF` (models.py:55-84) are NamedTuples of arrays plus the
model callables; here the callables must come from the compiled registry (`bayesianfiltering_b200.zoo`).
`NonlinearSSM.sample` (models.py:240-289) is the data generator; it runs on the GPU (`bf_ssm_sample`).
`NonlinearGaussianSSM` (models.py:86-157) and the distribution accessors of both classes are host-side carriers.
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


class MVN:
    """Host-side stand-in for the `MultivariateNormalFullCovariance` objects the reference's model classes hand out
    (models.py:126-157, :207-238): a carrier of (loc, covariance_matrix) with TFP's fp32 arithmetic -- `scale_tril =
    cholesky(cov)`, sample = `loc + L z`, log-prob through a triangular solve.  The filters never call it (their sampling
    and log-probabilities are inside the kernels); it exists so that code written against the model classes keeps working."""

    def __init__(self, loc, covariance_matrix):
        self.loc = np.asarray(loc, dtype=np.float32)
        self.covariance_matrix = np.asarray(covariance_matrix, dtype=np.float32)
        if self.covariance_matrix.shape != self.loc.shape[-1:] * 2:
            raise ValueError(f"covariance {self.covariance_matrix.shape} does not match loc {self.loc.shape}")

    def mean(self):
        return self.loc

    def covariance(self):
        return self.covariance_matrix

    def _scale_tril(self):
        return np.linalg.cholesky(self.covariance_matrix).astype(np.float32)

    def sample(self, sample_shape=(), seed=None):
        shape = (sample_shape,) if isinstance(sample_shape, (int, np.integer)) else tuple(sample_shape)
        z = np.random.default_rng(key_to_seed(seed)).standard_normal(shape + self.loc.shape).astype(np.float32)
        return self.loc + z @ self._scale_tril().T

    def log_prob(self, value):
        d = self.loc.shape[-1]
        L = self._scale_tril()
        y = np.linalg.solve(L, (np.asarray(value, dtype=np.float32) - self.loc)[..., None])[..., 0]
        return (-0.5 * np.sum(y * y, axis=-1) - np.sum(np.log(np.diag(L))) - 0.5 * d * np.log(2.0 * np.pi)).astype(np.float32)


def _registry_call(fn, state, noise, inputs):
    """Call a model function the way the reference's classes do.  A registry function (zoo) takes (x, noise, u) and is
    evaluated on the GPU; any other callable gets the reference's own argument lists."""
    if getattr(fn, "__bf_model__", None) is not None:
        return fn(state, noise, 0.0 if inputs is None else inputs)
    if noise is None:
        return fn(state) if inputs is None else fn(state, inputs)
    return fn(state, noise, inputs)


class NonlinearGaussianSSM(SSM):
    """Nonlinear model with ADDITIVE Gaussian noise, z_t ~ N(f(z_{t-1}, u_t), Q), y_t ~ N(h(z_t, u_t), R)
    (gaussfiltax/models.py:86-157).  Parameters travel in a `ParamsNLSSM`."""

    def __init__(self, state_dim: int, emission_dim: int, input_dim: int = 0):
        self.state_dim = state_dim
        self.emission_dim = emission_dim
        self.input_dim = 0          # sic: models.py:116

    @property
    def emission_shape(self):
        return (self.emission_dim,)

    @property
    def inputs_shape(self):
        return (self.input_dim,) if self.input_dim > 0 else None

    def initial_distribution(self, params, inputs=None):
        return MVN(params.initial_mean, params.initial_covariance)

    def transition_distribution(self, params, state, inputs=None):
        return MVN(_registry_call(params.dynamics_function, state, None, inputs), params.dynamics_noise_covariance)

    def emission_distribution(self, params, state, inputs=None):
        return MVN(_registry_call(params.emission_function, state, None, inputs), params.emission_noise_covariance)


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

    def initial_distribution(self, params, inputs=None):
        """models.py:207-212"""
        return MVN(params.initial_mean, params.initial_covariance)

    def transition_distribution(self, params, state, inputs=None):
        """models.py:214-225: N(f(state[, inputs]), Q) -- the reference calls `f` WITHOUT a noise argument here, which its own
        three-argument model functions do not accept; a registry function is evaluated at zero noise."""
        return MVN(_registry_call(params.dynamics_function, state, None, inputs), params.dynamics_noise_covariance)

    def emission_distribution(self, params, state, inputs=None):
        """models.py:227-238 (same remark)."""
        return MVN(_registry_call(params.emission_function, state, None, inputs), params.emission_noise_covariance)

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
