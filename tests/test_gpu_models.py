"""GPU tier: registry models (csrc/models.cuh) against the oracle's NumPy restatement and float64 finite
differences.  Tolerance: relative 1e-5 at the scale of each field (north_star)."""
import numpy as np
import pytest

from oracle import models as om
from tests.parity_cases import ALL_CASES, case, close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("u", [0.0, 1.0, 2.0])
def test_model_functions_match_oracle(name, u):
    c = case(name)
    m, z = c.omodel, c.zmodel_fn()
    rng = np.random.default_rng(11)
    n = 257
    x = (rng.normal(size=(n, m.dx)) + 0.3).astype(np.float32)
    q = (0.1 * rng.normal(size=(n, m.dq))).astype(np.float32)
    r = (0.1 * rng.normal(size=(n, m.dr))).astype(np.float32)
    close(z.evaluate(0, x, q, u), m.f(x, q, u), what="f")
    close(z.evaluate(1, x, r, u), m.h(x, r, u), what="h")
    q0 = np.zeros((n, m.dq), np.float32); r0 = np.zeros((n, m.dr), np.float32)
    x64 = x.astype(np.float64)
    close(z.evaluate(2, x, None, u), m.Fx(x64, q0.astype(np.float64), u), rtol=2e-5, atol_scale=2e-5, what="Fx")
    close(z.evaluate(3, x, None, u), m.Fq(x, q0, u), what="Fq")
    close(z.evaluate(4, x, None, u), m.Hx(x64, r0.astype(np.float64), u), rtol=2e-5, atol_scale=2e-5, what="Hx")
    fd = om.numerical_jacobian(lambda zz: m.f(zz, q0.astype(np.float64), u), x64)
    close(z.evaluate(2, x, None, u), fd, rtol=5e-5, atol_scale=5e-5, what="Fx vs finite differences")
    y = m.h(x[:1], r[:1], u)[0]
    lp = z.evaluate(5, x, y, u, R=c.R, r0=c.r0)
    ref = m.emission_log_prob(x64, y.astype(np.float64), u, c.r0, c.R)
    close(lp, ref, rtol=2e-5, atol_scale=1e-6, what="emission log-prob")


def test_tagged_callables_are_callable_like_the_reference_lambdas():
    c = case("ungm")
    z = c.zmodel_fn()
    x = np.array([0.7], np.float32)
    assert z.dynamics_function(x, np.array([0.1], np.float32), 2.0).shape == (1,)
    assert z.emission_function(x, 0.0, 0.0).shape == (1,)
    close(z.emission_function(x, 0.0, 0.0), [0.8 * 0.7])


def test_ssm_sample_matches_oracle_with_supplied_noise():
    """NonlinearSSM.sample (models.py:240-289) with the draws supplied."""
    from bayesianfiltering_b200.models import NonlinearSSM
    for name in ("ungm", "bot2", "lg42"):
        c = case(name)
        m = c.omodel
        T, B = 30, 3
        rng = np.random.default_rng(12)
        z0 = (c.m0 + 0.1 * rng.normal(size=(B, m.dx))).astype(np.float32)
        qn = (0.05 * rng.normal(size=(B, T, m.dq))).astype(np.float32)
        rn = (0.05 * rng.normal(size=(B, T, m.dr))).astype(np.float32)
        u = c.inputs(T)
        ssm = NonlinearSSM(m.dx, m.dq, m.dy, m.dr)
        s, e = ssm.sample(c.params_nlssm(), 0, T, inputs=u, num_trials=B, z0=z0, q_noise=qn, r_noise=rn)
        x = z0.copy()
        for t in range(T):
            ut = 0.0 if u is None else u[t]
            if t > 0:
                x = m.f(x, qn[:, t], ut)
            close(s[:, t], x, rtol=1e-4, atol_scale=1e-4, what=f"{name} state t={t}")
            close(e[:, t], m.h(s[:, t], rn[:, t], ut), what=f"{name} emission t={t}")
            x = s[:, t]          # teacher-force: chaotic maps amplify fp32 rounding


def test_ssm_sample_philox_statistics():
    from bayesianfiltering_b200.models import NonlinearSSM
    c = case("lg3")
    ssm = NonlinearSSM(3, 3, 3, 3)
    s, e = ssm.sample(c.params_nlssm(), 5, 50, num_trials=4000)
    assert s.shape == (4000, 50, 3) and e.shape == (4000, 50, 3)
    # stationary variance of x' = 0.8 x + q, q ~ N(0,1): 1/(1-0.64)
    v = s[:, -1].var(axis=0)
    np.testing.assert_allclose(v, 1.0 / 0.36, rtol=0.1)
    s2, _ = ssm.sample(c.params_nlssm(), 5, 50, num_trials=4000)
    assert np.array_equal(s, s2)                         # counter-based: same key, same draws


def test_model_classes_hand_out_distributions_of_registry_functions():
    """NonlinearGaussianSSM / NonlinearSSM accessors (models.py:126-157, :207-238) with a registry model: the mean is the
    model function evaluated on the GPU at zero noise."""
    from bayesianfiltering_b200.models import NonlinearGaussianSSM, NonlinearSSM
    c = case("l63")                     # additive noise of the state's dimension, as these accessors assume
    p, m = c.params_nlssm(), c.omodel
    x = np.array([1.0, -2.0, 20.0], np.float32)
    for ssm in (NonlinearGaussianSSM(3, 1), NonlinearSSM(3, 3, 1, 1)):
        t = ssm.transition_distribution(p, x, inputs=1.0)
        close(t.mean(), m.f(x[None], np.zeros((1, 3), np.float32), 1.0)[0], what="transition mean")
        assert np.array_equal(t.covariance(), c.Q)
        e = ssm.emission_distribution(p, x)
        close(e.mean(), m.h(x[None], np.zeros((1, 1), np.float32), 0.0)[0], what="emission mean")
        assert np.isfinite(e.log_prob(e.sample(seed=3)))
        assert np.array_equal(ssm.initial_distribution(p).mean(), c.m0)
