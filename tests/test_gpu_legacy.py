"""GPU tier: the legacy NumPy classes' call surface on the CUDA kernels, against
oracle/legacy.py -- the reference's loops restated in float64 with the random draws supplied.  The legacy classes are
fp64 and the kernels fp32, so the bar here is fp32 round-off on short runs (1e-4 at the scale of the field)."""
import numpy as np
import pytest

from oracle import legacy as oleg, models as om
from tests.parity_cases import close

pytestmark = pytest.mark.gpu


def _ssm(name):
    from bayesianfiltering_b200 import zoo
    from bayesianfiltering_b200.gaussfilt import SSM
    if name == "ungm":
        z, o, dx, dy = zoo.ungm(), om.UNGM(), 1, 1
        Q, R = np.array([[1.0]]), np.array([[1.0]])
    else:
        A, H = 0.8 * np.eye(3), 0.1 * np.eye(3) + 0.05
        z, o, dx, dy = zoo.linear_gaussian(A, H), om.LG(A, H), 3, 3
        Q, R = np.eye(3), 0.1 * np.eye(3)
    return SSM(dx, dy, np.zeros(dx), Q, np.zeros(dy), R, f=z.f, g=z.g), o, Q, R


def _data(ssm, o, Q, R, T, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(1, ssm.dx)); ys = []
    for t in range(T):
        x = o.f(x, rng.multivariate_normal(np.zeros(ssm.dx), Q)[None], 0.0)
        ys.append(o.h(x, rng.multivariate_normal(np.zeros(ssm.dy), R)[None], 0.0)[0])
    return np.stack(ys)


@pytest.mark.parametrize("name", ["ungm", "lg3"])
def test_legacy_gauss_sum_filters(name):
    from bayesianfiltering_b200.gaussfilt import EKF, GaussSumFilt as Bank
    from bayesianfiltering_b200.gausssumfilt import GaussSumFilt
    ssm, o, Q, R = _ssm(name)
    T, M = 12, 7
    ys = _data(ssm, o, Q, R, T, 3)
    m0, P0 = np.zeros(ssm.dx), np.eye(ssm.dx)
    init = m0 + np.random.default_rng(4).normal(size=(M, ssm.dx))
    for cls, ref in ((GaussSumFilt(ssm, M), oleg.gsf_legacy(o, ys, Q, R, P0, init, True, True)),
                     (Bank(EKF(ssm), M), oleg.gsf_legacy(o, ys, Q, R, P0, init, False, False))):
        got = cls.run(ys, m0, P0, init_means=init)
        assert got[0].shape == (T + 1, ssm.dx, M) and got[1].shape == (T + 1, ssm.dx, ssm.dx, M)
        assert got[2].shape == (T + 1, M) and got[3].shape == (T, ssm.dx)
        for g, r, what in zip(got, ref, ("means", "covs", "weights", "point_est")):
            close(g, r, rtol=2e-4, atol_scale=2e-4, what=f"{cls} {what}")
        assert cls.time > 0.0


def test_legacy_ekf_run_and_moment_approx():
    from bayesianfiltering_b200.gaussfilt import EKF
    ssm, o, Q, R = _ssm("lg3")
    T = 15
    ys = _data(ssm, o, Q, R, T, 5)
    m0, P0 = np.zeros(3), np.eye(3)
    ekf = EKF(ssm)
    ll, ms, Ps = ekf.run(ys, m0, P0)
    rll, rms, rPs = oleg.ekf_run(o, ys, Q, R, m0, P0)
    close(ll, rll, rtol=2e-4, atol_scale=2e-4); close(ms, rms, rtol=2e-4, atol_scale=2e-4); close(Ps, rPs, rtol=2e-4, atol_scale=2e-4)
    mean, var, cov = ekf.moment_approx(m0 + 0.3, P0, 'upd')
    assert mean.shape == (1, 3) and var.shape == (3, 3) and cov.shape == (3, 3)
    H = o.H
    close(var, R + H @ P0 @ H.T, rtol=1e-5, atol_scale=1e-5)


def test_legacy_bootstrap_pf_and_simulate():
    from bayesianfiltering_b200.particlefilt import BootstrapPF
    ssm, o, Q, R = _ssm("ungm")
    T, N = 10, 600
    ys = _data(ssm, o, Q, R, T, 6)
    rng = np.random.default_rng(7)
    x0 = rng.normal(size=(N, 1)); noise = rng.normal(size=(T, N, 1)); u = 1.0 - np.sort(rng.random((T, N)), axis=1)
    pf = BootstrapPF(ssm, N)
    got = pf.run(ys, np.zeros(1), np.eye(1), init_particles=x0, noise=noise.astype(np.float32), resample_uniforms=u)
    assert got.shape == (T + 1, N, 1) and np.array_equal(got[T], x0)
    ref = oleg.bootstrap_pf(o, ys, Q, R, x0.astype(np.float32).astype(np.float64), noise.astype(np.float32).astype(np.float64),
                            u.astype(np.float32).astype(np.float64))
    # the same ancestors up to the float32/float64 cdf resolution: compare the particle clouds through their moments
    for t in range(T):
        assert abs(got[t].mean() - ref[t].mean()) < 0.05 * (1.0 + ref[t].std())
        assert abs(got[t].std() - ref[t].std()) < 0.1 * (1.0 + ref[t].std())
    # default call (global NumPy generator), and the data generator
    assert pf.run(ys, np.zeros(1), np.eye(1)).shape == (T + 1, N, 1)
    xs, yy = ssm.simulate(50, np.zeros(1))
    assert xs.shape == (50, 1) and yy.shape == (50, 1) and np.all(np.isfinite(xs))
