"""Legacy Gaussian filters with the call surface of gaussfiltax/gaussfilt.py, running on libbfilt.so.

    SSM(dx, dy, c, Q, d, R, f, g)            gaussfilt.py:10-52   (simulate / propagate)
    EKF(ssm, order=2).run(ys, m0, P0)        gaussfilt.py:88-130 + :217-252  -> (ll, filtered_means, filtered_covs)
    GaussSumFilt(gauss_filt, num_models).run gaussfilt.py:311-361 -> (means [T+1,dx,M], covs [T+1,dx,dx,M], weights [T+1,M], point_est)

`f` and `g` must be the single-argument faces of a registry model (`zoo.ungm().f`, `.g`); additive noise
(`new_x = f(x) + N(c, Q)`), no inputs.  The legacy classes predict first and then update, store the initial condition
at index T of their output arrays and draw from NumPy's global generator; the draws can be supplied instead
(`init_means=`) for comparisons.  `EKF.moment_approx` is first order, as in the reference (its second-order terms are
commented out, gaussfilt.py:240-248), whatever `order` says.
"""
import time

import numpy as np
from numpy import random

from . import _lib, zoo
from .engine import GaussianSumFilterEngine


def _model_of(f, g):
    mf, mg = getattr(f, "__bf_model__", None), getattr(g, "__bf_model__", None)
    if mf is None or mg is None or mf is not mg:
        raise TypeError("f and g must be the .f / .g faces of one bayesianfiltering_b200.zoo model: arbitrary Python "
                        "callables cannot run inside the CUDA kernels and there is no CPU fallback")
    return mf


class SSM:
    """gaussfilt.py:10-52"""

    def __init__(self, dx, dy, c, Q, d, R, f=None, g=None):
        self.dx, self.dy, self.f, self.g, self.Q, self.R, self.c, self.d = dx, dy, f, g, Q, R, c, d

    def _host_model(self, m0=None, P0=None, with_bias=False):
        """The legacy FILTERS never use the noise means c, d (e.g. gaussfilt.py:241: mean_out = func(m)); only
        simulate/propagate do (gaussfilt.py:49-50)."""
        m = _model_of(self.f, self.g)
        if (m.dx, m.dy) != (self.dx, self.dy) or m.dq != m.dx:
            raise ValueError("registry model dims do not match this SSM (legacy classes need additive noise, dq = dx)")
        return m, m.host_model(q0=self.c if with_bias else None, Q=self.Q, r0=self.d if with_bias else None, R=self.R, m0=m0, P0=P0)

    def simulate(self, T, x0, key=None):
        """gaussfilt.py:21-42: xs[t] = f(xs[t-1]) + N(c,Q), ys[t] = g(xs[t]) + N(d,R), starting from x0."""
        from .models import NonlinearSSM, ParamsNLSSM
        self.T = T
        m, _ = self._host_model(with_bias=True)
        params = ParamsNLSSM(np.zeros(self.dx), np.eye(self.dx), m.dynamics_function, self.c, self.Q,
                             m.emission_function, self.d, self.R)
        seed = int(random.randint(0, 2 ** 31 - 1)) if key is None else key
        xs, ys = NonlinearSSM(m.dx, m.dq, m.dy, m.dr).sample(params, seed, T + 1, num_trials=1,
                                                             z0=np.asarray(x0, np.float32).reshape(1, self.dx))
        return xs[0, 1:].astype(np.float64), ys[0, 1:].astype(np.float64)

    def propagate(self, old_x):
        """gaussfilt.py:44-52"""
        xs, ys = self.simulate(1, old_x)
        return xs[0], ys[0]


class GaussFilt:
    """gaussfilt.py:55-130: a single Gaussian filter = the Gaussian-sum filter with one component."""

    def __init__(self, ssm):
        self.ssm = ssm
        self.f, self.g, self.Q, self.R, self.dx, self.dy = ssm.f, ssm.g, ssm.Q, ssm.R, ssm.dx, ssm.dy
        self.time = 0.0

    def moment_approx(self, m, P, kw):
        raise NotImplementedError

    def run(self, ys, m0, P0, verbose=False):
        """-> (ll [T], filtered_means [T,dx], filtered_covs [T,dx,dx])   (gaussfilt.py:88-130)"""
        tin = time.time()
        _, hm = self.ssm._host_model(m0, P0)
        eng = GaussianSumFilterEngine(hm, 1, 1, None, _lib.GSF_PREDICT_FIRST | _lib.GSF_NO_JITTER)
        out = eng.run_host(np.asarray(ys, np.float32).reshape(-1, self.dy), init_means=np.asarray(m0, np.float32).reshape(1, 1, self.dx),
                           init_covs=np.asarray(P0, np.float32).reshape(1, 1, self.dx, self.dx),
                           outputs=("means", "covariances", "log_likelihood"))
        eng.close()
        self.time = time.time() - tin
        return (out["log_likelihood"][0].astype(np.float64), out["means"][0, :, 0].astype(np.float64),
                out["covariances"][0, :, 0].astype(np.float64))


class EKF(GaussFilt):
    """gaussfilt.py:200-252"""

    def __init__(self, ssm, order=2):
        GaussFilt.__init__(self, ssm)
        self.order = order

    def __str__(self):
        return 'EKF'

    def moment_approx(self, m, P, kw):
        """First-order moments (gaussfilt.py:217-252): 'pred' -> (f(m), Q + F P F^T, P F^T); 'upd' with g, R."""
        model = _model_of(self.f, self.g)
        m = np.asarray(m, np.float32).reshape(1, self.dx)
        P = np.asarray(P, np.float64).reshape(self.dx, self.dx)
        if kw == 'pred':
            mean, J, cov, dim_out = model.evaluate(0, m), model.evaluate(2, m)[0].astype(np.float64), self.Q, self.dx
        elif kw == 'upd':
            mean, J, cov, dim_out = model.evaluate(1, m), model.evaluate(4, m)[0].astype(np.float64), self.R, self.dy
        else:
            raise ValueError(kw)
        var_out = np.asarray(cov, np.float64) + J @ P @ J.T
        return (np.reshape(mean.astype(np.float64), [1, dim_out]), np.reshape(var_out, [dim_out, dim_out]),
                np.reshape(P @ J.T, [self.dx, dim_out]))


class GaussSumFilt:
    """gaussfilt.py:311-361: a bank of `num_models` EKFs started at m0 + N(0, I), reweighted by their evidence."""
    _flags = _lib.GSF_PREDICT_FIRST | _lib.GSF_NO_JITTER

    def __init__(self, gauss_filt, num_models):
        if not isinstance(gauss_filt, EKF):
            raise NotImplementedError("only the EKF moment approximation runs on the GPU path (UKF/MCF/MCLAF are out of scope)")
        self.M = self.num_models = num_models
        self.dx, self.dy, self.gf = gauss_filt.dx, gauss_filt.dy, gauss_filt
        self.time = 0.0

    def __str__(self):
        return 'gf.GSF'

    def run(self, ys, m0, P0, verbose=False, init_means=None):
        tin = time.time()
        T = np.shape(ys)[0]
        if init_means is None:        # gaussfilt.py:335-338: m0 + N(0, I) from NumPy's global generator
            init_means = np.stack([m0 + random.multivariate_normal(np.zeros(self.dx), np.eye(self.dx)) for _ in range(self.M)])
        init_means = np.asarray(init_means, np.float32).reshape(self.M, self.dx)
        _, hm = self.gf.ssm._host_model(m0, P0)
        eng = GaussianSumFilterEngine(hm, self.M, 1, None, self._flags)
        out = eng.run_host(np.asarray(ys, np.float32).reshape(T, self.dy), init_means=init_means, outputs=("means", "covariances", "weights"))
        eng.close()
        means = np.zeros((T + 1, self.dx, self.M)); covs = np.zeros((T + 1, self.dx, self.dx, self.M)); weights = np.zeros((T + 1, self.M))
        means[:T] = np.transpose(out["means"][0], (0, 2, 1))
        covs[:T] = np.transpose(out["covariances"][0], (0, 2, 3, 1))
        weights[:T] = out["weights"][0]
        # initial conditions in the last entry (gaussfilt.py:332-338)
        means[T] = init_means.T; covs[T] = np.asarray(P0, np.float64).reshape(self.dx, self.dx)[:, :, None]; weights[T] = 1.0 / self.M
        # gaussfilt.py:358 (sic): np.sum without an axis -> one scalar broadcast over the state dimension
        point_est = np.zeros((T, self.dx))
        for t in range(T):
            point_est[t] = np.sum(np.multiply(means[t, :], weights[t]))
        self.time = time.time() - tin
        return means, covs, weights, point_est
