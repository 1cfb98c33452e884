"""Legacy Gaussian filters with the call surface of gaussfiltax/gaussfilt.py, running on libbfilt.so.

    SSM(dx, dy, c, Q, d, R, f, g)            gaussfilt.py:10-52   (simulate / propagate)
    EKF(ssm, order=2).run(ys, m0, P0)        gaussfilt.py:88-130 + :217-252  -> (ll, filtered_means, filtered_covs)
    UKF(ssm, alpha, beta, kappa).run(...)    gaussfilt.py:133-170   sigma points from chol(P), additive noise
    MCF(ssm, num_particles).run(...)         gaussfilt.py:173-198   Monte-Carlo moment approximation
    MCLAF                                    gaussfilt.py:255-308   raises: its moment_approx cannot run in the reference either
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

    def _legacy(self):
        """which moment approximation the kernel runs: None = EKF (Jacobians), else the `legacy=` tuple of the engine"""
        return None

    def run(self, ys, m0, P0, verbose=False, draws=None):
        """-> (ll [T], filtered_means [T,dx], filtered_covs [T,dx,dx])   (gaussfilt.py:88-130).
        draws [T,2,num_particles,dx] (MCF only): the Monte-Carlo draws as offsets x - m (else drawn in the kernel)."""
        tin = time.time()
        _, hm = self.ssm._host_model(m0, P0)
        eng = GaussianSumFilterEngine(hm, 1, 1, None, _lib.GSF_PREDICT_FIRST | _lib.GSF_NO_JITTER, legacy=self._legacy())
        kw = {} if draws is None else dict(mc_draws=np.asarray(draws, np.float32)[None, :, None])
        out = eng.run_host(np.asarray(ys, np.float32).reshape(-1, self.dy), init_means=np.asarray(m0, np.float32).reshape(1, 1, self.dx),
                           init_covs=np.asarray(P0, np.float32).reshape(1, 1, self.dx, self.dx),
                           outputs=("means", "covariances", "log_likelihood"), **kw)
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


class UKF(GaussFilt):
    """gaussfilt.py:133-170: unscented moment approximation, additive noise, 2 dx + 1 sigma points from chol(P)
    (utils.split_to_sigma_points, utils.py:52-65).  The kernels are float32: the class's default alpha = 1e-3 gives sigma-point
    weights of +-5e5 whose cancellation needs float64 -- pass a spread of order 1 (the reference's new-generation drivers use
    ParamsUKF(1, 0, 0)); small alpha raises."""

    def __init__(self, ssm, alpha=1e-3, beta=2, kappa=0):
        GaussFilt.__init__(self, ssm)
        self.alpha, self.beta, self.kappa = alpha, beta, kappa
        self.lamda = alpha ** 2 * (self.dx + kappa) - self.dx

    def __str__(self):
        return 'UKF'

    def _legacy(self):
        if self.alpha < 0.05:
            raise ValueError("UKF(alpha=%g): sigma-point weights of order 1/alpha^2 cancel catastrophically in float32; "
                             "use alpha of order 1 on the GPU path" % self.alpha)
        return ("ukf", self.alpha, self.beta, self.kappa)


class MCF(GaussFilt):
    """gaussfilt.py:173-198: Monte-Carlo moment approximation with `num_particles` draws of N(m, P) per prediction and per
    update (drawn in the kernel from Philox unless `draws=` is given to run)."""

    def __init__(self, ssm, num_particles, seed=None):
        GaussFilt.__init__(self, ssm)
        self.num_particles = num_particles
        self.seed = int(random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)

    def __str__(self):
        return 'MCF'

    def _legacy(self):
        return ("mcf", self.num_particles, self.seed)


class MCLAF(GaussFilt):
    """gaussfilt.py:255-308.  Not runnable in the reference: `moment_approx` calls `utils.sdp_opt` with 8 positional arguments
    (gaussfilt.py:286) where utils.py:110 takes 7 (TypeError; recorded in tests/golden/ref_legacy_gaussfilt.npz), so there is
    no behaviour to reproduce."""

    def __init__(self, ssm, num_particles):
        GaussFilt.__init__(self, ssm)
        self.num_particles = num_particles

    def __str__(self):
        return 'MCLAF'

    def run(self, ys, m0, P0, verbose=False):
        raise NotImplementedError("MCLAF.moment_approx raises TypeError in the reference itself (sdp_opt arity, gaussfilt.py:286 vs "
                                  "utils.py:110): nothing to match")


class GaussSumFilt:
    """gaussfilt.py:311-361: a bank of `num_models` Gaussian filters started at m0 + N(0, I), reweighted by their evidence."""
    _flags = _lib.GSF_PREDICT_FIRST | _lib.GSF_NO_JITTER

    def __init__(self, gauss_filt, num_models):
        if isinstance(gauss_filt, MCLAF) or not isinstance(gauss_filt, GaussFilt):
            raise NotImplementedError("the bank runs over EKF, UKF or MCF moment approximations")
        self.M = self.num_models = num_models
        self.dx, self.dy, self.gf = gauss_filt.dx, gauss_filt.dy, gauss_filt
        self.time = 0.0

    def __str__(self):
        return 'gf.GSF'

    def run(self, ys, m0, P0, verbose=False, init_means=None, draws=None):
        """draws [T,M,2,num_particles,dx] (bank of MCFs only): the Monte-Carlo draws as offsets x - m."""
        tin = time.time()
        T = np.shape(ys)[0]
        if init_means is None:        # gaussfilt.py:335-338: m0 + N(0, I) from NumPy's global generator
            init_means = np.stack([m0 + random.multivariate_normal(np.zeros(self.dx), np.eye(self.dx)) for _ in range(self.M)])
        init_means = np.asarray(init_means, np.float32).reshape(self.M, self.dx)
        _, hm = self.gf.ssm._host_model(m0, P0)
        eng = GaussianSumFilterEngine(hm, self.M, 1, None, self._flags, legacy=self.gf._legacy())
        kw = {} if draws is None else dict(mc_draws=np.asarray(draws, np.float32)[None])
        out = eng.run_host(np.asarray(ys, np.float32).reshape(T, self.dy), init_means=init_means, outputs=("means", "covariances", "weights"), **kw)
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
