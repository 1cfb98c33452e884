"""GPU tier: the CUDA kernels against the REFERENCE's own outputs.

`tests/golden/ref_*.npz` hold what the unmodified gaussfiltax source returned (run on the NumPy stand-in for jax/tfp,
oracle/make_ref_fixtures.py) together with every input and random draw it consumed.  The kernels get the same inputs
through the C ABI and must give the same outputs: step by step from the reference's own carry (teacher-forced, so that
one-ulp differences are not amplified by the nonlinear maps), relative 1e-5, widened only where the fp32 recursion
itself loses digits (`close_cond` measures that with the float64 oracle); whole trajectories on the well-conditioned
models.  Resampled ancestors: bit-exact on the kernel's own normalised weights, and equal to the reference's in all but
the few slots where the two weight vectors differ in their last bit.
"""
import os

import numpy as np
import pytest

from oracle import gsf as ogsf, pf as opf, scan as oscan
from tests.parity_cases import case, close, close_cond
from tests.test_ref_pin import ref, omodel, _inputs, FIELDS

pytestmark = pytest.mark.gpu
ALL_OUT = ("weights", "particles", "propagated", "norm_weights", "ancestors", "means", "log_likelihood", "ess", "resampled")


def zoo_model(name, d):
    from bayesianfiltering_b200 import zoo
    if name == "msv":
        th = d["theta"]; n = (th.size - 1) // 3
        return zoo.msv(th[:n], th[n:2 * n], float(th[2 * n]), th[2 * n + 1:])
    return case({"lg4": "lg42"}.get(name, name)).zmodel_fn()


def host_model(name, d, Q=None, R=None):
    z = zoo_model(name, d)
    return z.host_model(q0=d["q0"], Q=d["Q"] if Q is None else Q, r0=d["r0"], R=d["R"] if R is None else R, m0=d["m0"], P0=d["P0"])


def pf_teacher_forced(hm, om_, g, path, thr, res, ref_anc=None, ref_nw=None):
    """One kernel step from the fixture's carry at every t (fixture: weights [N,T], particles [N,T,dx] post-resample)."""
    from bayesianfiltering_b200.engine import ParticleFilterEngine
    ys = g["emissions"]; T = ys.shape[0]; N = g["x0"].shape[0]
    u = g["inputs"] if "inputs" in g.files else None
    W, X = g["weights"], g["particles"]
    eng = ParticleFilterEngine(hm, N, 1, thr, res, path, 0, 0)
    f8 = lambda v: np.asarray(v, dtype=np.float64)
    n_res = 0
    for t in range(T):
        x = g["x0"] if t == 0 else np.ascontiguousarray(X[:, t - 1])
        w = np.full(N, np.float32(1.0) / np.float32(N), np.float32) if t == 0 else np.ascontiguousarray(W[:, t - 1])
        ut = 0.0 if u is None else u[t]
        ru = g["res_u"][t:t + 1]
        out = eng.run_host(ys[t:t + 1], inputs=None if u is None else u[t:t + 1], init_particles=x, init_weights=w,
                           noise=g["noise"][t:t + 1], resample_u=ru, outputs=ALL_OUT)
        o = opf.pf_step(om_, x, w, ys[t], ut, g["noise"][t], g["res_u"][t], g["r0"], g["R"], thr, res)
        o8 = opf.pf_step(om_, f8(x), f8(w), f8(ys[t]), float(ut), f8(g["noise"][t]), g["res_u"][t], g["r0"], g["R"], thr, res)
        close(out["propagated"][0, 0], o["propagated"], what=f"propagated t={t}")
        close_cond(out["norm_weights"][0, 0], o["norm_weights"], o8["norm_weights"], what=f"norm_weights t={t}")
        flag = bool(out["resampled"][0, 0])
        if abs(float(o["ess"]) - thr * N) / (thr * N) > 1e-3:
            assert flag == bool(g["resampled"][t]), f"resample decision differs from the reference at t={t}"
        if flag:
            n_res += 1
            anc = oscan.resample_ancestors(out["norm_weights"][0, 0], g["res_u"][t], res)
            assert np.array_equal(out["ancestors"][0, 0], anc), f"ancestors not bit-exact on the kernel's own weights, t={t}"
            if ref_nw is not None:
                close_cond(out["norm_weights"][0, 0], ref_nw[t], o8["norm_weights"], what=f"norm_weights vs reference t={t}")
            ra = ref_anc[t] if ref_anc is not None else o["ancestors"]
            same = out["ancestors"][0, 0] == ra
            assert same.mean() >= 0.97, f"t={t}: {(~same).sum()} of {N} ancestors differ from the reference's"
            close_cond(out["particles"][0, 0][same], X[:, t][same], o8["propagated"][ra][same], what=f"particles vs fixture t={t}")
            assert np.all(out["weights"][0, 0] == np.float32(1.0) / np.float32(N))
        else:
            close_cond(out["weights"][0, 0], W[:, t], o8["weights"], what=f"weights vs fixture t={t}")
            close_cond(out["particles"][0, 0], X[:, t], o8["particles"], what=f"particles vs fixture t={t}")
    eng.close()
    return n_res


@pytest.mark.parametrize("path", ["resident", "streamed"])
@pytest.mark.parametrize("name", ["ungm", "bot2", "lg3", "bot1", "msv"])
def test_bootstrap_particle_filter_vs_reference(name, path):
    """inference.bootstrap_particle_filter (inference.py:1302-1380)"""
    d = ref(f"bpf_{name}")
    n = pf_teacher_forced(host_model(name, d), omodel(name, d), d, path, float(d["ess_threshold"]), "multinomial",
                          ref_anc=d["ancestors"], ref_nw=d["norm_weights_on_resample"])
    assert n > 0


def _gsf_engine_step(eng, d, t, M, fields=FIELDS):
    f = np.float32
    if t == 0:
        w = np.full((1, M), 1.0 / M, f); pm = d["means0"][None]; pP = np.broadcast_to(d["P0"], (1, M) + d["P0"].shape).copy()
    else:
        w = d["weights"][:, t - 1][None]; pm = d["predicted_means"][:, t - 1][None]; pP = d["predicted_covariances"][:, t - 1][None]
    u = d["inputs"][t:t + 1] if "inputs" in d.files else None
    kw = {}
    if d["Q"].ndim == 3:
        kw = dict(Q_t=d["Q"][t:t + 1], R_t=d["R"][t:t + 1])
    return eng.run_host(d["emissions"][t:t + 1], inputs=u, init_means=np.ascontiguousarray(pm), init_covs=np.ascontiguousarray(pP),
                        init_weights=np.ascontiguousarray(w), outputs=fields, **kw)


def _mixture_vs_reference(d, eng_factory, oracle_step, what):
    M = int(d["M"]); T = d["emissions"].shape[0]
    u = _inputs(d, T)
    eng = eng_factory()
    for t in range(T):
        out = _gsf_engine_step(eng, d, t, M)
        f8 = lambda a: np.asarray(a, np.float64)
        if t == 0:
            w = np.full(M, 1.0 / M); pm = f8(d["means0"]); pP = np.broadcast_to(f8(d["P0"]), (M,) + d["P0"].shape).copy()
        else:
            w, pm, pP = f8(d["weights"][:, t - 1]), f8(d["predicted_means"][:, t - 1]), f8(d["predicted_covariances"][:, t - 1])
        Q = f8(d["Q"][t] if d["Q"].ndim == 3 else d["Q"]); R = f8(d["R"][t] if d["R"].ndim == 3 else d["R"])
        o8 = oracle_step(w, pm, pP, f8(d["emissions"][t]), float(u[t]), Q, f8(d["q0"]), R, f8(d["r0"]))
        for k in FIELDS:
            close_cond(out[k][0, 0], d[k][:, t], o8[k], rtol=1e-5, what=f"{what} {k} t={t}")
    eng.close()


@pytest.mark.parametrize("name", ["ungm", "bot2", "bot1", "l63", "lg42", "lg42_tv", "msv"])
def test_gaussian_sum_filter_vs_reference(name):
    """inference.gaussian_sum_filter (inference.py:303-377), + whole trajectories where rounding is not amplified"""
    from bayesianfiltering_b200.engine import GaussianSumFilterEngine
    d = ref(f"gsf_{name}")
    base = name.replace("_tv", "")
    m = omodel(base, d)
    M = int(d["M"])
    tv = d["Q"].ndim == 3
    hm = host_model(base, d, Q=d["Q"][0] if tv else None, R=d["R"][0] if tv else None)
    _mixture_vs_reference(d, lambda: GaussianSumFilterEngine(hm, M, 1, 0),
                          lambda w, pm, pP, y, u, Q, q0, R, r0: ogsf.gsf_step(m, w, pm, pP, y, u, Q, q0, R, r0), f"gsf {name}")
    if name in ("ungm", "lg42", "lg42_tv", "msv", "l63"):
        eng = GaussianSumFilterEngine(hm, M, 1, 0)
        kw = dict(Q_t=d["Q"], R_t=d["R"]) if tv else {}
        out = eng.run_host(d["emissions"], inputs=d["inputs"] if "inputs" in d.files else None, init_means=d["means0"][None], outputs=FIELDS, **kw)
        eng.close()
        for k in FIELDS:
            close(np.swapaxes(out[k][0], 0, 1), d[k], rtol=2e-4, atol_scale=2e-5, what=f"gsf {name} trajectory {k}")


@pytest.mark.parametrize("name", ["ungm", "bot2", "l63", "lg42", "msv"])
def test_data_generator_vs_reference(name):
    """models.NonlinearSSM.sample (models.py:240-289) with the reference's own draws"""
    from bayesianfiltering_b200.models import NonlinearSSM, ParamsNLSSM
    d = ref(f"sample_{name}")
    z = zoo_model(name, ref("bpf_msv") if name == "msv" else d)
    T = d["emissions"].shape[0]
    params = ParamsNLSSM(d["m0"], d["P0"], z.dynamics_function, d["q0"], d["Q"], z.emission_function, d["r0"], d["R"])
    xs, ys = NonlinearSSM(z.dx, z.dq, z.dy, z.dr).sample(params, 0, T, inputs=d["inputs"] if "inputs" in d.files else None,
                                                         z0=d["z0"][None], q_noise=d["q_noise"][None], r_noise=d["r_noise"][None])
    close(xs, d["states"], rtol=1e-5, atol_scale=1e-6, what="states")
    close(ys, d["emissions"], rtol=1e-5, atol_scale=1e-6, what="emissions")


@pytest.mark.parametrize("name", ["ungm", "bot2", "l63", "lg42", "msv"])
def test_unscented_gaussian_sum_filter_vs_reference(name):
    """inference.unscented_gaussian_sum_filter (inference.py:379-456): sigma points of the noise-augmented state, sqrtm by
    Jacobi in fp32 against the reference's scipy.linalg.sqrtm"""
    from bayesianfiltering_b200.engine import GaussianSumFilterEngine
    from oracle import ugsf as ougsf
    d = ref(f"ugsf_{name}")
    m = omodel(name, d)
    M = int(d["M"])
    up = tuple(float(v) for v in d["uparams"])
    hm = host_model(name, d)
    _mixture_vs_reference(d, lambda: GaussianSumFilterEngine(hm, M, 1, 0, unscented=up),
                          lambda w, pm, pP, y, u, Q, q0, R, r0: ougsf.ugsf_step(m, w, pm, pP, y, u, Q, q0, R, r0, ougsf.ParamsUKF(*up)),
                          f"ugsf {name}")


def test_unscented_entry_point_signature_and_shapes():
    """unscented_gaussian_sum_filter(params, uparams, emissions, num_components, num_iter, inputs) -> PosteriorGaussianSumFiltered
    with [M,T,...] fields (inference.py:379-386, :451-456); on a linear model it is the Kalman filter for any spread."""
    from bayesianfiltering_b200 import inference as gf
    c = case("lg42")
    T, M = 12, 3
    _, ys = c.simulate(T, 4)
    means0 = np.tile(c.m0, (M, 1))
    post = gf.unscented_gaussian_sum_filter(c.params_nlssm(), gf.ParamsUKF(1, 0, 0), ys, M, 1, None, init_means=means0[None])
    assert post.weights.shape == (M, T) and post.means.shape == (M, T, 4) and post.covariances.shape == (M, T, 4, 4)
    assert post.predicted_means.shape == (M, T, 4) and post.predicted_covariances.shape == (M, T, 4, 4)
    ms, Ps, _ = ogsf.kalman_filter_exact(c.omodel.A, c.omodel.H, c.Q, c.R, c.m0, c.P0, ys)
    close(post.means[0], ms, rtol=2e-4, atol_scale=2e-5, what="UGSF means vs exact Kalman")
    close(post.covariances[0], Ps, rtol=2e-4, atol_scale=2e-5, what="UGSF covs vs exact Kalman")


def test_kalman_step_vs_reference():
    """inference._kalman_step (inference.py:107-120) through the package's face of the same name"""
    from bayesianfiltering_b200 import inference as gf
    d = ref("kalman_step")
    for name in ("ungm", "bot2", "l63", "lg42"):
        z = case(name).zmodel_fn()
        m = omodel(name)
        g = lambda k: d[f"{name}_{k}"]
        for k in range(g("m").shape[0]):
            ll, pm, pP = gf._kalman_step(g("m")[k], g("P")[k], z.dynamics_function, None, None, g("Q"), g("q0"), g("u")[k],
                                         z.emission_function, None, None, g("R"), g("r0"), g("y")[k])
            f8 = lambda a: np.asarray(a, np.float64)
            o8 = ogsf.kalman_step(m, f8(g("m")[k])[None], f8(g("P")[k])[None], f8(g("Q")), f8(g("q0")), f8(g("R")), f8(g("r0")),
                                  float(g("u")[k]), f8(g("y")[k]))
            close_cond(np.asarray(ll), g("ll")[k], o8[0][0], rtol=1e-5, what=f"{name} ll {k}")
            close_cond(pm, g("post_mean")[k], o8[1][0], rtol=1e-5, what=f"{name} mean {k}")
            close_cond(pP, g("post_cov")[k], o8[2][0], rtol=1e-5, what=f"{name} cov {k}")


@pytest.mark.parametrize("name", ["ungm", "bot2", "l63", "lg42"])
def test_speedy_agsf_vs_reference(name):
    """inference.speedy_augmented_gaussian_sum_filter (inference.py:621-812): every step from the reference's carry; the
    survivors' ancestors bit-exact on the kernel's own leaf weights and equal to the reference's"""
    from bayesianfiltering_b200.engine import AugmentedGaussianSumFilterEngine
    from oracle import agsf as oagsf
    d = ref(f"agsf_{name}")
    m = omodel(name)
    M, N1, N2 = (int(v) for v in d["num_components"])
    T = d["emissions"].shape[0]
    u = _inputs(d, T)
    eng = AugmentedGaussianSumFilterEngine(host_model(name, d), (M, N1, N2), 1, tuple(d["opt_args"]), 0, 0)
    outs = ("weights", "means", "covariances", "pre_weights", "updated_means", "ancestors")
    f8 = lambda a: np.asarray(a, np.float64)
    for t in range(T):
        fm = d["means0"] if t == 0 else d["means"][:, t - 1]
        fP = np.broadcast_to(d["P0"], (M,) + d["P0"].shape).copy() if t == 0 else d["covariances"][:, t - 1]
        out = eng.run_host(d["emissions"][t:t + 1], inputs=d["inputs"][t:t + 1] if "inputs" in d.files else None,
                           init_means=np.ascontiguousarray(fm)[None], init_covs=np.ascontiguousarray(fP)[None],
                           z_iid=d["z_iid"][t], s_iid=d["s_iid"][t], res_u=d["res_u"][t], outputs=outs)
        o8 = oagsf.agsf_step(m, f8(fm), f8(fP), np.full(M, 1.0 / M), f8(d["emissions"][t]), float(u[t]), f8(d["Q"]), f8(d["q0"]),
                             f8(d["R"]), f8(d["r0"]), f8(d["z_iid"][t]), f8(d["s_iid"][t]), f8(d["res_u"][t]), d["opt_args"])
        close_cond(out["pre_weights"][0, 0], d["pre_weights"][t].reshape(-1), o8["pre_weights"], rtol=1e-5, what=f"{name} pre_weights t={t}")
        close_cond(out["updated_means"][0, 0], d["updated_means"][t], o8["updated_means"], rtol=1e-5, what=f"{name} updated_means t={t}")
        anc = oscan.resample_ancestors(out["pre_weights"][0, 0], d["res_u"][t], "multinomial")
        assert np.array_equal(out["ancestors"][0, 0], anc), f"ancestors not bit-exact on the kernel's own weights, t={t}"
        same = out["ancestors"][0, 0] == d["res_idx"][t]
        assert same.sum() >= M - 1, f"t={t}: survivors {out['ancestors'][0, 0]} vs the reference's {d['res_idx'][t]}"
        idx = d["res_idx"][t]
        close_cond(out["means"][0, 0][same], d["means"][:, t][same], o8["updated_means"][idx][same], rtol=1e-5, what=f"{name} means t={t}")
        close_cond(out["covariances"][0, 0][same], d["covariances"][:, t][same], o8["updated_covs"][idx][same], rtol=1e-5, what=f"{name} covs t={t}")
        assert np.all(out["weights"][0, 0] == np.float32(1.0) / np.float32(M))
    eng.close()


def test_speedy_agsf_entry_point_signature_and_shapes():
    """speedy_augmented_gaussian_sum_filter(params, emissions, num_components, rng_key, num_iter, opt_args, inputs) ->
    (PosteriorGaussianSumFiltered(weights, means, covariances), aux_outputs)   (inference.py:621-629, :803-812)"""
    from bayesianfiltering_b200 import inference as gf
    from bayesianfiltering_b200.random import PRNGKey
    d = ref("agsf_bot2")
    c = case("bot2")
    nc = [int(v) for v in d["num_components"]]
    post, aux = gf.speedy_augmented_gaussian_sum_filter(c.params_nlssm(), d["emissions"], nc, rng_key=PRNGKey(3), opt_args=tuple(d["opt_args"]),
                                                        inputs=d["inputs"], init_means=d["means0"][None], z_iid=d["z_iid"][0], s_iid=d["s_iid"][0],
                                                        resample_uniforms=d["res_u"][0])
    T, M, L = d["emissions"].shape[0], nc[0], nc[0] * nc[1] * nc[2]
    assert post.weights.shape == (M, T) and post.means.shape == (M, T, 4) and post.covariances.shape == (M, T, 4, 4)
    assert post.predicted_means is None
    assert aux["pre_weights"].shape == (T, L) and aux["updated_means"].shape == (T, L, 4)
    # the reference's own first step (the later ones depend on which leaves survived, to the last bit of the weights)
    close(post.means[:, 0], d["means"][:, 0], rtol=1e-4, atol_scale=1e-5, what="first step means")
    assert np.all(np.isfinite(post.means))


def test_optimal_resampling_vs_reference():
    """utils.optimal_resampling (utils.py:216-244)"""
    from bayesianfiltering_b200 import utils
    d = ref("optimal_resampling")
    k = 0
    while f"w{k}" in d.files:
        idx, w = utils.optimal_resampling(d[f"w{k}"], int(d[f"N{k}"]), 0, uniforms=d[f"u{k}"])
        assert np.array_equal(idx, d[f"idx{k}"]), (k, idx, d[f"idx{k}"])
        close(w, d[f"weights{k}"], rtol=1e-5, atol_scale=1e-6, what=f"optimal_resampling weights {k}")
        k += 1
    assert k >= 3


# ------------------------------------------------------------------------------------------ legacy NumPy classes
def _legacy_ssm(name, d):
    """the single-argument faces of the registry models the legacy fixtures were generated with"""
    from bayesianfiltering_b200 import zoo
    from bayesianfiltering_b200.gaussfilt import SSM
    g = lambda k: d[f"{name}_{k}"]
    z = zoo.lorenz63() if name == "l63" else zoo.ungm(c=0.0)
    dx, dy = g("Q").shape[0], g("R").shape[0]
    return SSM(dx, dy, np.zeros(dx), g("Q"), np.zeros(dy), g("R"), f=z.f, g=z.g)


@pytest.mark.parametrize("name", ["l63", "ungm"])
@pytest.mark.parametrize("kind", ["EKF", "UKF", "MCF"])
def test_legacy_gaussfilt_classes_vs_reference(name, kind):
    """gaussfilt.{EKF, UKF, MCF}(ssm).run (gaussfilt.py:88-252): the float64 NumPy reference against the float32 kernels, on
    15 steps; the Monte-Carlo filter gets the reference's own draws"""
    from bayesianfiltering_b200 import gaussfilt as gfm
    d = ref("legacy_gaussfilt")
    g = lambda k: d[f"{name}_{k}"]
    ssm = _legacy_ssm(name, d)
    if kind == "EKF":
        ll, ms, Ps = gfm.EKF(ssm, order=1).run(g("ys"), g("m0"), g("P0"))
    elif kind == "UKF":
        ll, ms, Ps = gfm.UKF(ssm, alpha=1.0, beta=2.0, kappa=0.5).run(g("ys"), g("m0"), g("P0"))
    else:
        ll, ms, Ps = gfm.MCF(ssm, g("MCF_draws").shape[2]).run(g("ys"), g("m0"), g("P0"), draws=g("MCF_draws"))
    tol = 5e-3 if (kind == "MCF" and name == "l63") else 5e-4     # fp32 against fp64 over 15 steps; the 48-draw sample covariance is ill-conditioned on L63
    close(ms, g(f"{kind}_means"), rtol=tol, atol_scale=tol, what=f"{kind} means")
    close(Ps, g(f"{kind}_covs"), rtol=10 * tol, atol_scale=tol, what=f"{kind} covs")
    close(ll, g(f"{kind}_ll"), rtol=10 * tol, atol_scale=tol, what=f"{kind} ll")


def test_legacy_bank_of_ukfs_vs_reference_and_mclaf_raises():
    """gaussfilt.GaussSumFilt(UKF(ssm), M).run (gaussfilt.py:325-361) on the scalar model; MCLAF has no behaviour to match"""
    from bayesianfiltering_b200 import gaussfilt as gfm
    d = ref("legacy_bank")
    g = lambda k: d[f"ungm_{k}"]
    ssm = _legacy_ssm("ungm", d)
    M = g("UKF_init_means").shape[0]
    out = gfm.GaussSumFilt(gfm.UKF(ssm, alpha=1.0, beta=2.0, kappa=0.5), M).run(g("ys"), g("m0"), g("P0"), init_means=g("UKF_init_means"))
    for a, k in zip(out, ("means", "covs", "weights", "point")):
        close(a, g(f"UKF_{k}"), rtol=2e-3, atol_scale=5e-4, what=f"bank of UKFs {k}")
    with pytest.raises(NotImplementedError):
        gfm.MCLAF(ssm, 8).run(g("ys"), g("m0"), g("P0"))
    with pytest.raises(ValueError):
        gfm.UKF(ssm).run(g("ys"), g("m0"), g("P0"))            # default alpha = 1e-3: refused in float32
