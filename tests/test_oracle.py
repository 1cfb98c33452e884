"""CPU tier: the oracle against independent known answers (the reference ships no golden vectors --
SURVEY.md section 4 -- so these pins are ours) and against the committed fixtures in tests/golden/."""
import os

import numpy as np
import pytest
import scipy.stats

from oracle import gsf as ogsf, models as om, pf as opf, scan as oscan
from tests.parity_cases import ALL_CASES, case

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 31, 33, 100, 1000, 1024, 4097])
def test_cumsum_recursive_equals_closed_form(n):
    rng = np.random.default_rng(n)
    x = rng.random(n).astype(np.float32)
    x /= x.sum(dtype=np.float32)
    a = oscan.cumsum_associative(x)
    assert np.array_equal(a, oscan.cumsum_closed_form(x))
    np.testing.assert_allclose(a, np.cumsum(x.astype(np.float64)), rtol=0, atol=4e-7)


def test_cumsum_differs_from_sequential_in_last_bits():
    rng = np.random.default_rng(0)
    x = rng.random(4097).astype(np.float32); x /= x.sum(dtype=np.float32)
    assert not np.array_equal(oscan.cumsum_associative(x), np.cumsum(x))   # the order matters: that is the point


def test_bisect_equals_numpy_on_sorted():
    rng = np.random.default_rng(1)
    for n in [1, 2, 5, 64, 1000, 4097]:
        c = np.sort(rng.random(n).astype(np.float32))
        r = np.concatenate([rng.random(3000).astype(np.float32), c[: min(n, 50)], [0.0, 1.0]]).astype(np.float32)
        assert np.array_equal(oscan.searchsorted_bisect(c, r), np.searchsorted(c, r, "left"))


@pytest.mark.parametrize("kind", ["weights", "noisy", "garbage", "flat"])
def test_fixed_bisection_is_a_lower_bound_on_effective_thresholds(kind):
    """The identity the large-N CUDA path rests on: jnp.searchsorted's fixed bisection on ANY array equals
    numpy's searchsorted on the monotone array oracle.scan.effective_thresholds_pow2 derives from it."""
    rng = np.random.default_rng({"weights": 0, "noisy": 1, "garbage": 2, "flat": 3}[kind])
    for n in [1, 2, 4, 8, 64, 512, 4096]:
        if kind == "weights":
            w = rng.random(n).astype(np.float32) ** 3; w /= w.sum(dtype=np.float32)
            cdf = oscan.cumsum_associative(w)
        elif kind == "noisy":
            cdf = np.sort(rng.random(n).astype(np.float32))
            idx = rng.integers(0, n, size=max(1, n // 4)); cdf[idx] += rng.normal(0, 0.02, size=idx.size).astype(np.float32)
        elif kind == "garbage":
            cdf = rng.random(n).astype(np.float32)
        else:
            w = np.zeros(n, np.float32); w[rng.integers(n)] = 1.0
            cdf = oscan.cumsum_associative(w) + rng.integers(-2, 3, size=n).astype(np.float32) * np.float32(6e-8)
        tau = oscan.effective_thresholds_pow2(cdf)
        assert np.all(np.diff(tau) >= 0)
        r = np.concatenate([rng.random(2000).astype(np.float32), cdf, np.nextafter(cdf, np.float32(2)),
                            np.nextafter(cdf, np.float32(-1)), [np.float32(-1), np.float32(2)]]).astype(np.float32)
        assert np.array_equal(oscan.searchsorted_bisect(cdf, r), np.searchsorted(tau, r, "left"))
        if kind == "weights":
            assert np.mean(tau == cdf) > 0.9          # a last-bit repair, not a different cdf


def test_resample_invariants():
    rng = np.random.default_rng(2)
    n = 3000
    w = rng.random(n).astype(np.float32) ** 4
    w[::7] = 0.0
    w /= w.sum(dtype=np.float32)
    a = oscan.resample_ancestors(w, rng.random(n, dtype=np.float32), "multinomial")
    assert a.min() >= 0 and a.max() < n
    assert not np.any(w[a] == 0.0)                      # zero-weight particles are never selected (side='left')
    s = oscan.resample_ancestors(w, np.float32(0.37), "systematic")
    assert np.all(np.diff(s) >= 0)                      # sorted positions -> monotone ancestors
    counts = np.bincount(s, minlength=n)
    assert counts.sum() == n and np.all(np.abs(counts - n * w.astype(np.float64)) < 1.0 + 1e-3 * n * w)


def test_mvn_log_prob_against_scipy():
    rng = np.random.default_rng(3)
    for k in (1, 2, 5):
        A = rng.normal(size=(k, k)); cov = A @ A.T + k * np.eye(k)
        loc = rng.normal(size=(7, k)); y = rng.normal(size=k)
        ref = np.array([scipy.stats.multivariate_normal(l, cov).logpdf(y) for l in loc])
        np.testing.assert_allclose(om.mvn_log_prob(loc, cov, y), ref, rtol=1e-12)


@pytest.mark.parametrize("name", ALL_CASES)
def test_jacobians_against_finite_differences(name):
    c = case(name)
    m = c.omodel
    rng = np.random.default_rng(4)
    x = rng.normal(size=(6, m.dx)) + 0.5
    q = np.zeros((6, m.dq)); r = np.zeros((6, m.dr))
    for u in (0.0, 1.0, 2.0, 0.3):
        np.testing.assert_allclose(m.Fx(x, q, u), om.numerical_jacobian(lambda z: m.f(z, q, u), x), rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(m.Hx(x, r, u), om.numerical_jacobian(lambda z: m.h(z, r, u), x), rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(m.Fq(x, q, u), om.numerical_jacobian(lambda z: m.f(x, z, u), q), rtol=2e-6, atol=2e-7)


def test_gsf_single_component_is_the_kalman_filter():
    """Linear-Gaussian model: gaussian_sum_filter(M=1) == closed-form KF up to the reference's psd_solve
    jitter (utils.py:258 adds 1e-6 to every element of S)."""
    c = case("lg42")
    T = 60
    _, ys = c.simulate(T, 5)
    km, kP, kll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
    o = ogsf.gaussian_sum_filter(c.omodel, ys.astype(np.float64), None, c.m0[None].astype(np.float64), c.P0, c.Q, c.q0, c.R, c.r0)
    np.testing.assert_allclose(o["means"][0], km, rtol=0, atol=5e-5)
    np.testing.assert_allclose(o["covariances"][0], kP, rtol=0, atol=5e-5)
    np.testing.assert_allclose(o["loglik_inc"], kll, rtol=0, atol=5e-5)
    assert abs(o["weights"] - 1.0).max() == 0.0


def test_pf_loglik_converges_to_kalman():
    """Config 5 in miniature: BPF marginal log-likelihood -> exact KF log-likelihood as N grows."""
    c = case("lg3")
    T = 25
    _, ys = c.simulate(T, 6)
    _, _, kll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
    errs = []
    for N in (200, 20000):
        x0, noise, ru = c.pf_randomness(T, N, 7)
        o = opf.bootstrap_particle_filter(c.omodel, ys, None, x0, noise, ru, c.r0, c.R, 0.5, keep_trajectory=False)
        errs.append(abs(o["marginal_loglik"] - kll.sum()))
    assert errs[1] < 0.5 and errs[1] < errs[0] + 0.2


def test_pf_outputs_shapes_and_invariants():
    c = case("ungm")
    T, N = 12, 300
    _, ys = c.simulate(T, 8)
    x0, noise, ru = c.pf_randomness(T, N, 9)
    o = opf.bootstrap_particle_filter(c.omodel, ys, c.inputs(T), x0, noise, ru, c.r0, c.R, 0.5)
    assert o["weights"].shape == (N, T) and o["particles"].shape == (N, T, 1)       # inference.py:1378
    np.testing.assert_allclose(o["weights"].sum(axis=0), 1.0, atol=1e-5)
    assert o["resampled"].sum() > 0


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLD) if f.endswith(".npz") and not f.startswith("ref_")) if os.path.isdir(GOLD) else [])
def test_oracle_reproduces_committed_golden(fname):
    """tests/golden/*.npz other than ref_* were written by tests/golden/make_golden.py from this oracle; they freeze it.
    (ref_*.npz come from the reference itself: tests/test_ref_pin.py)"""
    from tests.golden.make_golden import regenerate
    g = np.load(os.path.join(GOLD, fname))
    fresh = regenerate(fname[:-4])
    for k in g.files:
        if g[k].dtype.kind == "f":
            np.testing.assert_allclose(fresh[k], g[k], rtol=1e-6, atol=1e-7, err_msg=f"{fname}:{k}")
        else:
            assert np.array_equal(fresh[k], g[k]), f"{fname}:{k}"


def test_oracle_time_varying_covariances():
    """`_get_params` (inference.py:21): covariances with a leading T axis are indexed by the step; constant matrices
    repeated T times give the same filter; GSF(M=1) on a linear-Gaussian model is the time-varying Kalman filter."""
    from oracle import gsf as ogsf
    from tests.parity_cases import case
    c = case("lg3")
    T = 15
    _, ys = c.simulate(T, 5)
    f8 = lambda v: np.asarray(v, dtype=np.float64)
    a = ogsf.gaussian_sum_filter(c.omodel, f8(ys), None, f8(c.m0[None]), f8(c.P0), f8(c.Q), f8(c.q0), f8(c.R), f8(c.r0))
    b = ogsf.gaussian_sum_filter(c.omodel, f8(ys), None, f8(c.m0[None]), f8(c.P0), np.broadcast_to(f8(c.Q), (T, 3, 3)), f8(c.q0),
                                 np.broadcast_to(f8(c.R), (T, 3, 3)), f8(c.r0))
    for k in ("means", "covariances", "loglik_inc"):
        assert np.array_equal(a[k], b[k]), k
    rng = np.random.default_rng(0)
    Qt = rng.uniform(0.5, 2.0, T)[:, None, None] * f8(c.Q)[None]
    Rt = rng.uniform(0.5, 2.0, T)[:, None, None] * f8(c.R)[None]
    g = ogsf.gaussian_sum_filter(c.omodel, f8(ys), None, f8(c.m0[None]), f8(c.P0), Qt, f8(c.q0), Rt, f8(c.r0))
    ms, Ps, ll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, Qt, Rt, c.m0, c.P0, ys)
    np.testing.assert_allclose(g["means"][0], ms, rtol=1e-4, atol=1e-5)      # psd_solve's 1e-6 jitter is the difference
    np.testing.assert_allclose(g["loglik_inc"], ll, rtol=1e-4, atol=1e-5)
    assert np.max(np.abs(g["means"][0] - a["means"][0])) > 1e-3               # and it is not the constant filter


def test_unscented_oracle_sigma_points_and_linear_exactness():
    """oracle/ugsf.py (oracle first): the sigma points reproduce (m, P); on a linear-Gaussian model
    the unscented transform is exact, so UGSF(M=1) is the Kalman filter -- for ParamsUKF(1, 0, 0) (the setting of
    BOT_Experiment_script.py:108) and, in float64, for the default alpha = 1e-3 whose weights are ~1e6."""
    from oracle import gsf as ogsf, ugsf as ou
    from tests.parity_cases import case
    rng = np.random.default_rng(0)
    A = rng.standard_normal((3, 5, 5))
    P = A @ np.swapaxes(A, 1, 2) + 0.1 * np.eye(5)
    m = rng.standard_normal((3, 5))
    for lam in (0.0, 1.5, -4.99):
        sp = ou.get_sigma_points(m, P, lam)
        assert sp.shape == (3, 10, 5)
        np.testing.assert_allclose(sp.mean(axis=1), m, atol=1e-10)
        d = sp - m[:, None, :]
        np.testing.assert_allclose(np.einsum('mij,mik->mjk', d, d) / (2 * (5 + lam)), P, rtol=1e-9, atol=1e-10)
    root = ou.sym_sqrtm(P)
    np.testing.assert_allclose(root @ root, P, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(root, np.swapaxes(root, 1, 2), atol=1e-12)
    for name in ("lg3", "lg42"):
        c = case(name)
        T = 25
        _, ys = c.simulate(T, 7)
        f8 = lambda v: np.asarray(v, dtype=np.float64)
        ms, Ps, ll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
        for up in (ou.ParamsUKF(1.0, 0.0, 0.0), ou.ParamsUKF()):
            g = ou.unscented_gaussian_sum_filter(c.omodel, f8(ys), None, f8(c.m0[None]), f8(c.P0), f8(c.Q), f8(c.q0), f8(c.R),
                                                 f8(c.r0), up)
            # psd_solve's 1e-6 jitter (utils.py:258) is the only difference from the closed form
            np.testing.assert_allclose(g["means"][0], ms, rtol=2e-4, atol=2e-5, err_msg=f"{name} {up}")
            np.testing.assert_allclose(g["covariances"][0], Ps, rtol=2e-4, atol=2e-5, err_msg=f"{name} {up}")
            np.testing.assert_allclose(g["loglik_inc"], ll, rtol=2e-4, atol=2e-5, err_msg=f"{name} {up}")
            assert np.all(g["weights"] == 1.0)
    # a nonlinear model, several components: finite, normalised weights, symmetric covariances
    c = case("ungm")
    T, M = 15, 6
    _, ys = c.simulate(T, 8)
    g = ou.unscented_gaussian_sum_filter(c.omodel, ys.astype(np.float64), c.inputs(T), c.gsf_means0(M, 9).astype(np.float64),
                                         c.P0.astype(np.float64), c.Q.astype(np.float64), c.q0.astype(np.float64),
                                         c.R.astype(np.float64), c.r0.astype(np.float64), ou.ParamsUKF(1.0, 0.0, 0.0))
    assert np.all(np.isfinite(g["means"])) and np.all(np.isfinite(g["loglik_inc"]))
    np.testing.assert_allclose(g["weights"].sum(axis=0), 1.0, rtol=1e-12)


@pytest.mark.parametrize("kind", ["flat", "lognormal", "sparse", "peaked"])
@pytest.mark.parametrize("n", [8, 256, 4096])
def test_cell_table_search_is_the_fixed_bisection(n, kind):
    """The large-N kernel's multinomial shortcut (two adjacent entries of a table of lower bounds at the cell edges + a search
    inside the cell, csrc/pf_fused.cuh) restated in NumPy: the reference's ancestors for every uniform, including the ends."""
    rng = np.random.default_rng(n + len(kind))
    w = {"flat": lambda: rng.random(n, dtype=np.float32),
         "lognormal": lambda: np.exp(6 * rng.standard_normal(n)).astype(np.float32),
         "sparse": lambda: np.where(rng.random(n) < 0.7, 0, rng.random(n, dtype=np.float32)).astype(np.float32) + np.float32(1e-3) * (np.arange(n) == n // 3),
         "peaked": lambda: rng.random(n, dtype=np.float32) ** 20}[kind]()
    w = (w / w.sum(dtype=np.float32)).astype(np.float32)
    for u in (rng.random(n, dtype=np.float32),                                                   # any float32 uniform
              rng.integers(0, 1 << 24, n).astype(np.float32) / np.float32(1 << 24),              # the kernel's 24-bit uniforms: r is a cell edge
              np.concatenate([[0.0, np.float32(1) - np.float32(2.0 ** -24), 0.5], rng.random(n - 3)]).astype(np.float32)):
        assert np.array_equal(oscan.cell_table_ancestors(w, u), oscan.resample_ancestors(w, u, "multinomial"))
