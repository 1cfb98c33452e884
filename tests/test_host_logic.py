"""CPU tier: host-side mirror of the reference's call surface (no GPU, no kernels)."""
import inspect

import numpy as np
import pytest

from bayesianfiltering_b200 import containers, inference, models, zoo
from bayesianfiltering_b200.random import PRNGKey, key_to_seed, split


def test_param_containers_have_the_reference_fields():
    # gaussfiltax/models.py:43-50 and :73-80
    base = ("initial_mean", "initial_covariance", "dynamics_function", "dynamics_noise_bias",
            "dynamics_noise_covariance", "emission_function", "emission_noise_bias", "emission_noise_covariance")
    assert models.ParamsNLSSM._fields == base
    assert models.ParamsBPF._fields == base + ("emission_distribution_log_prob",)
    assert hasattr(models.ParamsBPF, "sample_dynamics_distribution")
    # inference.py:35-39
    assert inference.PosteriorGaussianSumFiltered._fields == ("weights", "means", "covariances", "predicted_means",
                                                              "predicted_covariances")
    assert containers.GaussianComponent._fields == ("mean", "covariance", "weight")
    assert containers.GaussianSum._fields == ("means", "covariances", "weights")


def test_entry_point_signatures_match_the_reference():
    # inference.py:1302-1309
    p = list(inspect.signature(inference.bootstrap_particle_filter).parameters.values())
    assert [q.name for q in p[:6]] == ["params", "emissions", "num_particles", "key", "inputs", "ess_threshold"]
    assert p[5].default == 0.5 and p[4].default is None
    assert all(q.kind is inspect.Parameter.KEYWORD_ONLY for q in p[6:])
    # inference.py:303-309
    g = list(inspect.signature(inference.gaussian_sum_filter).parameters.values())
    assert [q.name for q in g[:5]] == ["params", "emissions", "num_components", "num_iter", "inputs"]
    assert g[2].default == 1 and g[3].default == 1 and g[4].default is None
    assert all(q.kind is inspect.Parameter.KEYWORD_ONLY for q in g[5:])
    s = inspect.signature(models.NonlinearSSM.__init__)
    assert list(s.parameters)[1:6] == ["state_dim", "state_noise_dim", "emission_dim", "emission_noise_dim", "input_dim"]


def test_untagged_callables_raise_no_fallback():
    f = lambda x, q, u: x + q
    p = models.ParamsNLSSM(np.zeros(1), np.eye(1), f, np.zeros(1), np.eye(1), f, np.zeros(1), np.eye(1))
    with pytest.raises(TypeError, match="no CPU fallback"):
        inference.gaussian_sum_filter(p, np.zeros((3, 1), np.float32), 2)
    with pytest.raises(TypeError, match="no CPU fallback"):
        inference.bootstrap_particle_filter(p, np.zeros((3, 1), np.float32), 10)


def test_mixed_registry_models_raise():
    a, b = zoo.ungm(), zoo.ungm()
    p = models.ParamsNLSSM(np.zeros(1), np.eye(1), a.dynamics_function, np.zeros(1), np.eye(1), b.emission_function,
                           np.zeros(1), np.eye(1))
    with pytest.raises(ValueError):
        zoo.resolve(p)


def test_host_model_arrays_and_shape_checks():
    m = zoo.bot(emission_dim=2)
    hm = m.host_model(Q=1e-5 * np.eye(2), R=25e-6 * np.eye(2), m0=[-0.05, 0.001, 0.7, -0.05], P0=np.diag([0.1, 0.005, 0.1, 0.01]))
    assert (hm.dx, hm.dq, hm.dy, hm.dr) == (4, 2, 2, 2) and hm.desc.n_theta == 3
    with pytest.raises(ValueError):
        m.host_model(Q=np.eye(3))
    with pytest.raises(ValueError):
        zoo.bot(emission_dim=3)


def test_keys():
    k = PRNGKey(7)
    assert k.dtype == np.uint32 and k.shape == (2,) and key_to_seed(k) == 7
    s = split(k, 3)
    assert s.shape == (3, 2) and len({key_to_seed(x) for x in s}) == 3
    assert key_to_seed(None) == 0 and key_to_seed(5) == 5


def test_time_varying_params_are_rejected_loudly():
    m = zoo.ungm()
    p = models.ParamsNLSSM(np.zeros(1), np.eye(1), m.dynamics_function, np.zeros(1), np.ones((5, 1, 1)),
                           m.emission_function, np.zeros(1), np.eye(1))
    with pytest.raises(NotImplementedError):
        zoo.host_model_from_params(p)


def test_time_varying_covariances_are_routed_like_get_params():
    """`_get_params` (inference.py:21): a covariance with a leading T axis is per-step; the description handed to the
    library holds its first matrix, the [T,d,d] arrays travel separately; biases are never time-indexed; the particle
    filter entry point refuses (the reference fetches the per-step matrices there and then does not use them)."""
    import numpy as np
    from bayesianfiltering_b200 import zoo
    from tests.parity_cases import case
    c = case("lg3")
    T = 7
    p = c.params_nlssm()
    Qt = np.stack([(1.0 + t) * c.Q for t in range(T)])
    Rt = np.stack([(2.0 + t) * c.R for t in range(T)])
    m, hm, q, r = zoo.host_model_from_params(p._replace(dynamics_noise_covariance=Qt, emission_noise_covariance=Rt), time_varying=True)
    assert q.shape == (T, 3, 3) and r.shape == (T, 3, 3)
    assert np.allclose(np.ctypeslib.as_array(hm.desc.Q, (9,)).reshape(3, 3), Qt[0])
    assert np.allclose(np.ctypeslib.as_array(hm.desc.R, (9,)).reshape(3, 3), Rt[0])
    m, hm, q, r = zoo.host_model_from_params(p, time_varying=True)
    assert q is None and r is None
    with pytest.raises(NotImplementedError):
        zoo.host_model_from_params(p._replace(dynamics_noise_covariance=Qt))
    with pytest.raises(NotImplementedError):
        zoo.host_model_from_params(p._replace(dynamics_noise_bias=np.zeros((T, 3), np.float32)), time_varying=True)


def test_bench_cpu_arm_and_yardsticks():
    """bench.py's host-side pieces: the CPU arm's workload (oracle/bench_workload.py) is the `bot2` parity case with the same
    draws; the Kalman yardstick of configs[4] equals the oracle's exact filter; the product never imports oracle/ or tests/."""
    import ast, os
    import bench
    from oracle import bench_workload, gsf as ogsf, pf as opf
    from tests.parity_cases import case
    secs, ll = bench_workload.run_bot(3, 1 << 10, 6)
    c = case("bot2")
    _, ys = c.simulate(6, 1003)
    x0, noise, ru = c.pf_randomness(6, 1 << 10, 2003, "systematic")
    o = opf.bootstrap_particle_filter(c.omodel, ys, c.inputs(6), x0, noise, ru, c.r0, c.R, 1.0, "systematic", keep_trajectory=False)
    assert secs > 0 and ll == float(o["marginal_loglik"])
    c = case("lg8")
    _, ys = c.simulate(25, 9)
    _, _, inc = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
    mine = bench._kalman_loglik(np.asarray(c.omodel.A, np.float64), np.asarray(c.omodel.H, np.float64), c.Q, c.R, c.m0, c.P0, ys)
    assert abs(mine - float(inc.sum())) < 1e-9 * abs(mine)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "bayesianfiltering_b200")
    for fn in [os.path.join(pkg, f) for f in os.listdir(pkg) if f.endswith(".py")] + [os.path.join(root, "bench.py")]:
        tree = ast.parse(open(fn).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import): mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom): mods = [node.module or ""]
            for m in mods:
                assert not m.startswith("tests"), f"{fn} imports {m}"
                if fn.startswith(pkg): assert not m.startswith("oracle"), f"{fn} imports {m}"


def test_model_classes_and_their_distribution_accessors():
    # gaussfiltax/models.py:86-157 (NonlinearGaussianSSM) and :207-238 (NonlinearSSM): host-side carriers
    from scipy.stats import multivariate_normal
    g = models.NonlinearGaussianSSM(2, 1)
    assert list(inspect.signature(models.NonlinearGaussianSSM.__init__).parameters)[1:] == ["state_dim", "emission_dim", "input_dim"]
    assert g.emission_shape == (1,) and g.inputs_shape is None and g.input_dim == 0
    P0 = np.array([[2.0, 0.3], [0.3, 1.0]], np.float32)
    p = models.ParamsNLSSM(np.array([1.0, 2.0]), P0, lambda z, *u: 2.0 * np.asarray(z), np.zeros(2), 0.5 * np.eye(2),
                           lambda z, *u: np.asarray(z)[:1], np.zeros(1), np.eye(1))
    d0 = g.initial_distribution(p)
    x = np.array([[0.5, 1.5], [2.0, 2.0]], np.float32)
    np.testing.assert_allclose(d0.log_prob(x), multivariate_normal([1, 2], P0).logpdf(x), rtol=2e-6)
    assert d0.sample(seed=PRNGKey(3)).shape == (2,) and d0.sample((4, 3), seed=PRNGKey(3)).shape == (4, 3, 2)
    assert np.array_equal(d0.sample(5, seed=PRNGKey(3)), d0.sample((5,), seed=PRNGKey(3)))          # a key fixes the draw
    s = d0.sample(200000, seed=PRNGKey(1))
    np.testing.assert_allclose(s.mean(0), [1, 2], atol=2e-2); np.testing.assert_allclose(np.cov(s.T), P0, atol=3e-2)
    t = g.transition_distribution(p, np.ones(2))
    assert np.array_equal(t.mean(), [2.0, 2.0]) and np.array_equal(t.covariance(), 0.5 * np.eye(2, dtype=np.float32))
    assert np.array_equal(g.emission_distribution(p, np.ones(2), inputs=0.0).mean(), [1.0])
    n = models.NonlinearSSM(2, 2, 1, 1)
    assert np.array_equal(n.initial_distribution(p).mean(), [1.0, 2.0]) and np.array_equal(n.transition_distribution(p, np.ones(2)).mean(), [2.0, 2.0])
    with pytest.raises(ValueError):
        models.MVN(np.zeros(2), np.eye(3))


def test_generic_ssm_sample_follows_the_accessors():
    # gaussfiltax/ssm.py:169-211, inherited by NonlinearGaussianSSM: x' = 0.8 x + N(0, I), y = x[0] + N(0, 0.01)
    g = models.NonlinearGaussianSSM(2, 1)
    p = models.ParamsNLSSM(np.zeros(2), np.eye(2), lambda z, *u: 0.8 * np.asarray(z), np.zeros(2), np.eye(2),
                           lambda z, *u: np.asarray(z)[:1], np.zeros(1), 0.01 * np.eye(1))
    s, e = g.sample(p, PRNGKey(0), 4000)
    assert s.shape == (4000, 2) and e.shape == (4000, 1) and s.dtype == np.float32
    np.testing.assert_allclose(s[200:].var(0), 1.0 / 0.36, rtol=0.2)
    np.testing.assert_allclose((e[:, 0] - s[:, 0]).std(), 0.1, rtol=0.1)
    s2, e2 = g.sample(p, PRNGKey(0), 4000)
    assert np.array_equal(s, s2) and np.array_equal(e, e2)                       # a key fixes the trajectory
    assert not np.array_equal(s[:50], g.sample(p, PRNGKey(1), 50)[0])
    seen = []
    pu = p._replace(dynamics_function=lambda z, u: (seen.append(float(u)), np.asarray(z) + u)[1])
    g.sample(pu, PRNGKey(2), 4, inputs=np.array([10.0, 20.0, 30.0, 40.0]))
    assert seen == [20.0, 30.0, 40.0]                                            # step t uses inputs[t]; the first state has no transition
    assert g.log_prior(p) == 0.0


def test_inference_helper_faces():
    # gaussfiltax/inference.py:21-26
    import scipy.stats
    Q = np.arange(12.0).reshape(3, 2, 2)
    assert np.array_equal(inference._get_params(Q, 2, 1), Q[1]) and np.array_equal(inference._get_params(Q[0], 2, 1), Q[0])
    assert inference._process_input(None, 4).shape == (4, 1) and inference._process_input("u", 4) == "u"
    S = np.array([[0.5, 0.1], [0.1, 0.3]])
    np.testing.assert_allclose(inference._MVN_log_prob(np.array([1.0, -1.0]), S, np.array([0.7, -0.2])),
                               scipy.stats.multivariate_normal([1.0, -1.0], S).logpdf([0.7, -0.2]), rtol=2e-6)
    assert inference._MVN_log_prob(np.zeros(1), np.eye(1), 0.3).shape == ()
    out = inference.swap_axes_on_values({"w": np.zeros((5, 3)), "x": np.zeros((5, 3, 2))})
    assert out["w"].shape == (3, 5) and out["x"].shape == (3, 5, 2)
