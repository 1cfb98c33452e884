"""GPU tier: Gaussian-sum (EKF mixture) filter (bf_gsf_* through the C ABI) against the oracle's restatement of
gaussfiltax/inference.py:303-377 (`_condition_on` :72-105, `_predict` :51-70, psd_solve utils.py:256-259).

Tolerance: relative 1e-5 in fp32 at the scale of each field (north_star).  `P - K S K^T` cancels, and the EKF
maps amplify rounding, so -- as for the particle filter -- the 1e-5 bar is checked per step from the oracle's
carry (`test_gsf_stepwise_parity`), whole trajectories at 1e-5 on linear-Gaussian models (where the filter is
also checked against the closed-form Kalman filter) and at a looser stated tolerance elsewhere.
"""
import os

import numpy as np
import pytest

from oracle import gsf as ogsf
from tests.parity_cases import case, close, close_cond

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FIELDS = ("weights", "means", "covariances", "predicted_means", "predicted_covariances")
ALL_OUT = FIELDS + ("log_likelihood", "mixture_mean")


def _engine(c, M, B=1):
    from bayesianfiltering_b200.engine import GaussianSumFilterEngine
    return GaussianSumFilterEngine(c.host_model(), M, B, 0)


@pytest.mark.parametrize("name,T,M", [
    ("ungm", 30, 100),      # BASELINE config 2 shape: M=100 components
    ("ungm_q10", 20, 33),
    ("bot1", 20, 8), ("bot2", 20, 5), ("l63", 20, 6), ("lg3", 20, 4), ("lg42", 20, 7), ("lg8", 12, 3),
    ("sin1d", 20, 5), ("sin3d", 4, 40),      # sin3d: the fp32 EKF loses positive definiteness at t=4
    ("l96s", 12, 9), ("l96", 6, 5),          # large-state kernel (gsf_large.cuh): Lorenz-96 d=8 and BASELINE configs[3]'s d=40
    ("l96_6", 8, 13), ("l96_10", 8, 13),     # d = 2 mod 4: every other warp's workspace would be misaligned without the padding
    ("l96", 10, 256),                        # configs[3]'s mixture size: every warp of k_gsf_large<40> loops over ~43 components
])
def test_gsf_stepwise_parity(name, T, M):
    """Each step from the ORACLE's carry (weights, predicted means, predicted covariances): one `_step`
    (inference.py:333-365) of the kernel against one `_step` of the oracle at 1e-5."""
    c = case(name)
    _, ys = c.simulate(T, 31)
    u = c.inputs(T)
    means0 = c.gsf_means0(M, 32)
    w = np.full(M, np.float32(1.0) / np.float32(M), np.float32)
    pm, pP = means0, np.broadcast_to(c.P0, (M,) + c.P0.shape).copy()
    eng = _engine(c, M)
    for t in range(T):
        ut = 0.0 if u is None else u[t]
        o = ogsf.gsf_step(c.omodel, w, pm, pP, ys[t], ut, c.Q, c.q0, c.R, c.r0)
        out = eng.run_host(ys[t:t + 1], inputs=None if u is None else u[t:t + 1], init_means=pm, init_covs=pP,
                           init_weights=w, outputs=ALL_OUT)
        f8 = lambda v: np.asarray(v, dtype=np.float64)
        o8 = ogsf.gsf_step(c.omodel, f8(w), f8(pm), f8(pP), f8(ys[t]), ut, f8(c.Q), f8(c.q0), f8(c.R), f8(c.r0))
        for k in FIELDS:
            close_cond(out[k][0, 0], o[k], o8[k], what=f"{name} {k} t={t}")
        # the increment is max + log(sum): its rounding is set by the size of the per-component log-likelihoods
        ll_scale = max(1.0, float(np.max(np.abs(o["lls"]))))
        ll_tol = 1e-5 * ll_scale + 8 * abs(float(o["loglik_inc"]) - float(o8["loglik_inc"]))
        assert abs(float(out["log_likelihood"][0, 0]) - float(o["loglik_inc"])) <= ll_tol, f"{name} loglik t={t}"
        close(out["mixture_mean"][0, 0], np.sum(o["weights"][:, None].astype(np.float64) * o["means"], axis=0),
              rtol=2e-5, atol_scale=2e-5, what=f"{name} mixture mean t={t}")
        w, pm, pP = o["weights"], o["predicted_means"], o["predicted_covariances"]
    eng.close()


@pytest.mark.parametrize("gold", ["gsf_ungm", "gsf_bot1", "gsf_lg42", "gsf_l63"])
def test_gsf_golden_fixtures(gold):
    """Whole trajectories against the committed oracle outputs, reference axis order [M,T,...]."""
    from tests.golden.make_golden import GSF_CASES
    from bayesianfiltering_b200 import inference as gf
    cname, T, M = GSF_CASES[gold]
    g = np.load(os.path.join(GOLD, gold + ".npz"))
    c = case(cname)
    u = g["inputs"] if "inputs" in g.files else None
    post, extras = gf.gaussian_sum_filter(c.params_nlssm(), g["emissions"], M, 1, u, init_means=g["means0"],
                                          outputs=ALL_OUT, return_extras=True)
    tol = 1e-5 if cname.startswith("lg") else 5e-3
    for k in FIELDS:
        a = getattr(post, k)
        assert a.shape == g[k].shape, k
        close(a, g[k], rtol=tol, atol_scale=tol, what=f"{gold} {k}")
    close(extras["log_likelihood_increments"], g["loglik_inc"], rtol=tol, atol_scale=tol)


def test_gsf_single_component_on_linear_model_is_the_kalman_filter():
    """SURVEY 8(c) analytical pin: GSF(M=1) on a linear-Gaussian model equals the closed-form Kalman filter
    (means, covariances and log-likelihood)."""
    for name in ("lg3", "lg42", "lg8"):
        c = case(name)
        T = 40
        _, ys = c.simulate(T, 41)
        eng = _engine(c, 1)
        out = eng.run_host(ys, init_means=c.m0[None], outputs=ALL_OUT)
        eng.close()
        ms, Ps, ll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
        # psd_solve's +1e-6 on every element of S is part of the reference (utils.py:258): tolerance covers it
        close(out["means"][0, :, 0], ms, rtol=2e-4, atol_scale=2e-4, what=name + " means")
        close(out["covariances"][0, :, 0], Ps, rtol=2e-4, atol_scale=2e-4, what=name + " covs")
        close(out["log_likelihood"][0], ll, rtol=2e-4, atol_scale=2e-4, what=name + " loglik")
        assert np.all(out["weights"] == 1.0)


def _tv(c, T, seed):
    """Time-varying noise covariances [T,d,d]: the case's matrices scaled by a per-step factor in [0.5, 2]."""
    rng = np.random.default_rng(seed)
    fq = rng.uniform(0.5, 2.0, T).astype(np.float32)
    fr = rng.uniform(0.5, 2.0, T).astype(np.float32)
    return fq[:, None, None] * c.Q[None], fr[:, None, None] * c.R[None]


def test_gsf_time_varying_covariances_single_component_is_the_kalman_filter():
    """`_get_params` (inference.py:21, :337-340): a noise covariance with a leading T axis is indexed by the step.
    GSF(M=1) on a linear-Gaussian model with time-varying Q_t, R_t equals the closed-form time-varying Kalman filter."""
    for name in ("lg3", "lg42"):
        c = case(name)
        T = 30
        _, ys = c.simulate(T, 43)
        Qt, Rt = _tv(c, T, 44)
        eng = _engine(c, 1)
        out = eng.run_host(ys, init_means=c.m0[None], outputs=ALL_OUT, Q_t=Qt, R_t=Rt)
        const = eng.run_host(ys, init_means=c.m0[None], outputs=ALL_OUT)
        tiled = eng.run_host(ys, init_means=c.m0[None], outputs=ALL_OUT, Q_t=np.broadcast_to(c.Q, (T,) + c.Q.shape),
                             R_t=np.broadcast_to(c.R, (T,) + c.R.shape))
        eng.close()
        ms, Ps, ll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, Qt, Rt, c.m0, c.P0, ys)
        close(out["means"][0, :, 0], ms, rtol=2e-4, atol_scale=2e-4, what=name + " means")
        close(out["covariances"][0, :, 0], Ps, rtol=2e-4, atol_scale=2e-4, what=name + " covs")
        close(out["log_likelihood"][0], ll, rtol=2e-4, atol_scale=2e-4, what=name + " loglik")
        for k in ALL_OUT:                                   # constant matrices repeated T times: the same filter, bit for bit
            assert np.array_equal(tiled[k], const[k]), k


@pytest.mark.parametrize("name,T,M", [("ungm", 12, 20), ("bot2", 10, 5), ("l96s", 8, 4)])
def test_gsf_time_varying_covariances_stepwise(name, T, M):
    """One call with Q_t / R_t over T steps against the oracle's whole run with the same per-step matrices, compared
    step by step from the kernel's OWN carry (teacher-forced, as everywhere else)."""
    c = case(name)
    _, ys = c.simulate(T, 45)
    u = c.inputs(T)
    Qt, Rt = _tv(c, T, 46)
    means0 = c.gsf_means0(M, 47)
    eng = _engine(c, M)
    out = eng.run_host(ys, inputs=u, init_means=means0, outputs=ALL_OUT, Q_t=Qt, R_t=Rt)
    eng.close()
    w = np.full(M, np.float32(1.0) / np.float32(M), np.float32)
    pm, pP = means0, np.broadcast_to(c.P0, (M,) + c.P0.shape).copy()
    f8 = lambda v: np.asarray(v, dtype=np.float64)
    for t in range(T):
        ut = 0.0 if u is None else u[t]
        o = ogsf.gsf_step(c.omodel, w, pm, pP, ys[t], ut, Qt[t], c.q0, Rt[t], c.r0)
        o8 = ogsf.gsf_step(c.omodel, f8(w), f8(pm), f8(pP), f8(ys[t]), ut, f8(Qt[t]), f8(c.q0), f8(Rt[t]), f8(c.r0))
        for k in FIELDS:
            close_cond(out[k][0, t], o[k], o8[k], what=f"{name} {k} t={t}")
        w, pm, pP = out["weights"][0, t], out["predicted_means"][0, t], out["predicted_covariances"][0, t]


def test_gsf_time_varying_covariances_through_the_reference_entry_point():
    from bayesianfiltering_b200 import inference as gf
    from bayesianfiltering_b200.models import ParamsNLSSM
    c = case("lg3")
    T = 12
    _, ys = c.simulate(T, 48)
    Qt, Rt = _tv(c, T, 49)
    m = c.zmodel_fn()
    params = ParamsNLSSM(c.m0, c.P0, m.dynamics_function, c.q0, Qt, m.emission_function, c.r0, Rt)
    post, ex = gf.gaussian_sum_filter(params, ys, 1, init_means=c.m0[None], return_extras=True,
                                      outputs=FIELDS + ("log_likelihood",))
    ms, Ps, ll = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, Qt, Rt, c.m0, c.P0, ys)
    close(post.means[0], ms, rtol=2e-4, atol_scale=2e-4, what="means")
    close(ex["log_likelihood_increments"], ll, rtol=2e-4, atol_scale=2e-4, what="loglik")
    with pytest.raises(NotImplementedError):              # the particle filter: not indexed by the reference either
        from bayesianfiltering_b200.models import ParamsBPF
        pb = ParamsBPF(c.m0, c.P0, m.dynamics_function, c.q0, Qt, m.emission_function, c.r0, c.R, m.emission_log_prob)
        gf.bootstrap_particle_filter(pb, ys, 64)


def test_gsf_trial_batch_and_chunked_resume():
    c = case("ungm")
    T, M, B = 25, 100, 5
    ys = np.stack([c.simulate(T, 50 + b)[1] for b in range(B)])
    u = c.inputs(T)
    means0 = np.stack([c.gsf_means0(M, 60 + b) for b in range(B)])
    eng = _engine(c, M, B)
    full = eng.run_host(ys, inputs=u, init_means=means0, outputs=ALL_OUT)
    k = 11
    a = eng.run_host(ys[:, :k], inputs=u[:k], init_means=means0, outputs=ALL_OUT)
    b = eng.run_host(ys[:, k:], inputs=u[k:], outputs=ALL_OUT, resume=True)
    eng.close()
    for name in ALL_OUT:
        assert np.array_equal(np.concatenate([a[name], b[name]], axis=1), full[name]), name
    for tr in range(B):
        one = _engine(c, M, 1)
        o = one.run_host(ys[tr], inputs=u, init_means=means0[tr], outputs=ALL_OUT)
        one.close()
        for name in ALL_OUT:
            assert np.array_equal(o[name][0], full[name][tr]), (name, tr)
    np.testing.assert_allclose(full["weights"].sum(axis=-1), 1.0, rtol=1e-5)


def test_gsf_device_tensor_mode_and_shared_emissions():
    import torch
    c = case("bot2")
    T, M, B = 15, 6, 4
    _, ys = c.simulate(T, 70)
    u = c.inputs(T)
    means0 = np.stack([c.gsf_means0(M, 80 + b) for b in range(B)])
    eng = _engine(c, M, B)
    host = eng.run_host(ys, inputs=u, init_means=means0, outputs=ALL_OUT)       # [T,dy] shared by all trials
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dev = eng.run(cu(ys), inputs=cu(u), init_means=cu(means0), outputs=ALL_OUT)
    torch.cuda.synchronize()
    for k in ALL_OUT:
        assert np.array_equal(dev[k].cpu().numpy(), host[k]), k
    assert eng.launches() == 2
    eng.close()


def test_gsf_lorenz96_whole_run_batch_and_resume():
    """BASELINE configs[3] shape at test size: Lorenz-96 d=40, M components x B trials, against the oracle over a
    whole run (the map is only mildly expanding at dt = 0.01), trial batching and chunked resume bit for bit."""
    c = case("l96")
    T, M, B = 8, 6, 3
    ys = np.stack([c.simulate(T, 300 + b)[1] for b in range(B)])
    means0 = np.stack([c.gsf_means0(M, 310 + b) for b in range(B)])
    eng = _engine(c, M, B)
    full = eng.run_host(ys, init_means=means0, outputs=ALL_OUT)
    a = eng.run_host(ys[:, :3], init_means=means0, outputs=ALL_OUT)
    b2 = eng.run_host(ys[:, 3:], outputs=ALL_OUT, resume=True)
    eng.close()
    for name in ALL_OUT:
        assert np.array_equal(np.concatenate([a[name], b2[name]], axis=1), full[name]), name
    for b in range(B):
        ref = ogsf.gaussian_sum_filter(c.omodel, ys[b], None, means0[b], c.P0, c.Q, c.q0, c.R, c.r0)
        ref8 = ogsf.gaussian_sum_filter(c.omodel, ys[b].astype(np.float64), None, means0[b].astype(np.float64), c.P0.astype(np.float64),
                                        c.Q.astype(np.float64), c.q0.astype(np.float64), c.R.astype(np.float64), c.r0.astype(np.float64))
        for k in FIELDS:
            close_cond(np.swapaxes(full[k][b], 0, 1), ref[k], ref8[k], rtol=2e-5, k=16.0, what=f"l96 {k} trial {b}")
        ll_scale = max(1.0, float(np.max(np.abs(ref["loglik_inc"]))))
        assert np.max(np.abs(full["log_likelihood"][b] - ref["loglik_inc"])) <= 2e-5 * ll_scale + 16 * np.max(np.abs(ref["loglik_inc"] - ref8["loglik_inc"]))
    np.testing.assert_allclose(full["weights"].sum(axis=-1), 1.0, rtol=1e-5)
    with pytest.raises(ValueError):                      # full noise covariances are not supported on the large-state path
        from bayesianfiltering_b200 import zoo
        from bayesianfiltering_b200.engine import GaussianSumFilterEngine
        GaussianSumFilterEngine(zoo.lorenz96(8).host_model(Q=np.eye(8, dtype=np.float32) + 0.1), 2, 1, 0)


def test_monte_carlo_trial_driver_single_process():
    """montecarlo.*_trials without a process group: the reference drivers' Nsim loop (BOT_Experiment_script.py:89) as one call."""
    from bayesianfiltering_b200 import montecarlo as mc
    c = case("lg3")
    T, B = 20, 6
    sims = [c.simulate(T, 400 + b) for b in range(B)]
    xs = np.stack([s[0] for s in sims]); ys = np.stack([s[1] for s in sims])
    g = mc.gaussian_sum_filter_trials(c.params_nlssm(), ys, 4, true_states=xs, key=3)
    p = mc.particle_filter_trials(c.params_bpf(), ys, 4096, true_states=xs, key=3)
    for out in (g, p):
        assert out["means"].shape == (B, T, 3) and out["rmse"].shape == (B,) and out["marginal_log_likelihood"].shape == (B,)
        assert np.all(np.isfinite(out["rmse"]))
    # a linear-Gaussian model: both filters estimate the same posterior mean
    assert np.mean(np.abs(g["means"] - p["means"])) < 0.15
    m, (lo, hi) = mc.bootstrap_ci(p["rmse"])
    assert lo <= m <= hi


def test_gsf_errors_are_loud():
    from bayesianfiltering_b200.engine import GaussianSumFilterEngine
    c = case("ungm")
    with pytest.raises(ValueError):
        GaussianSumFilterEngine(c.host_model(), 0, 1)
    with pytest.raises(ValueError):
        GaussianSumFilterEngine(c.host_model(), 5000, 1)
    eng = _engine(c, 4)
    with pytest.raises(ValueError):
        eng.run_host(np.zeros((3, 1), np.float32))                     # init_means missing
    with pytest.raises(ValueError):
        eng.run_host(np.zeros((3, 1), np.float32), resume=True)         # no carry yet
    eng.close()


def test_gsf_nan_emission_propagates_like_the_reference():
    """The reference lets NaN propagate; no silent clamping."""
    c = case("ungm")
    T, M = 6, 8
    _, ys = c.simulate(T, 90)
    ys = ys.copy(); ys[3] = np.nan
    eng = _engine(c, M)
    out = eng.run_host(ys, inputs=c.inputs(T), init_means=c.gsf_means0(M, 91), outputs=("weights", "means"))
    eng.close()
    assert np.all(np.isfinite(out["weights"][0, :3])) and np.all(np.isnan(out["weights"][0, 3:]))
