"""CPU tier: the oracle is PINNED to the reference.

`tests/golden/ref_*.npz` were produced by `oracle/make_ref_fixtures.py`, which executes the UNMODIFIED reference source
(`/root/reference/gaussfiltax`) on top of a NumPy stand-in for jax / tensorflow_probability (`oracle/refshim.py`; its header
lists what is executed as written and which third-party semantics are restated).  Each test feeds the fixture's recorded
inputs -- emissions, inputs, every random draw -- to the oracle's restatement and requires the reference's outputs.

Tolerances: fp32 filters are compared step by step from the reference's own previous state (`close_cond`: relative 1e-5,
widened only where the fp32 recursion itself loses digits, measured against the same restatement in float64), and as whole
trajectories on the well-conditioned models; the float64 legacy classes at 1e-9.
"""
import os

import numpy as np
import pytest

from oracle import agsf as oagsf, gsf as ogsf, legacy as oleg, models as om, pf as opf, scan as oscan, ugsf as ougsf
from tests.parity_cases import case, close, close_cond

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("weights", "means", "covariances", "predicted_means", "predicted_covariances")


def ref(name):
    return np.load(os.path.join(GOLD, f"ref_{name}.npz"))


def omodel(name, d=None):
    if name == "msv":
        th = d["theta"]; n = (th.size - 1) // 3
        return om.MSV(th[:n], th[n:2 * n], th[2 * n], th[2 * n + 1:])
    return case({"lg4": "lg42"}.get(name, name)).omodel


def _inputs(d, T):
    return d["inputs"] if "inputs" in d.files else np.zeros(T, np.float32)


def _cast(d, keys, dt):
    return [np.asarray(d[k], dtype=dt) for k in keys]


# ------------------------------------------------------------------------------------------ data generator
@pytest.mark.parametrize("name", ["ungm", "bot2", "l63", "lg42", "msv"])
def test_sample_matches_reference(name):
    """models.NonlinearSSM.sample (models.py:240-289): z_0 ~ N(m0,P0), y_0 = h(z_0, r_0, u_0); then z_t = f(z_{t-1}, q_t, u_t)."""
    d = ref(f"sample_{name}")
    m = omodel(name, ref("bpf_msv") if name == "msv" else None)
    T = d["emissions"].shape[0]
    u = _inputs(d, T)
    x = d["z0"][None].astype(np.float32)
    xs, ys = [], []
    for t in range(T):
        if t > 0:
            x = m.f(x, d["q_noise"][t][None], u[t])
        xs.append(x[0]); ys.append(m.h(x, d["r_noise"][t][None], u[t])[0])
    close(np.stack(xs), d["states"], rtol=1e-5, atol_scale=1e-6, what="states")
    close(np.stack(ys), d["emissions"], rtol=1e-5, atol_scale=1e-6, what="emissions")


# ------------------------------------------------------------------------------------------ Gaussian-sum filters
def _mixture_stepwise(d, step_fn, M, what):
    """teacher-forced: every step restarts from the REFERENCE's carry (weights, predicted moments of step t-1)"""
    T = d["emissions"].shape[0]
    u = _inputs(d, T)
    worst = 0.0
    for t in range(T):
        outs = {}
        for dt in (np.float32, np.float64):
            f = lambda a: np.asarray(a, dtype=dt)
            if t == 0:
                w = np.full(M, 1.0 / M, dtype=dt); pm = f(d["means0"]); pP = np.broadcast_to(f(d["P0"]), (M,) + d["P0"].shape).copy()
            else:
                w, pm, pP = f(d["weights"][:, t - 1]), f(d["predicted_means"][:, t - 1]), f(d["predicted_covariances"][:, t - 1])
            Q = f(d["Q"][t] if d["Q"].ndim == 3 else d["Q"]); R = f(d["R"][t] if d["R"].ndim == 3 else d["R"])
            outs[dt] = step_fn(w, pm, pP, f(d["emissions"][t]), dt(u[t]), Q, f(d["q0"]), R, f(d["r0"]))
        for k in FIELDS:
            close_cond(d[k][:, t], outs[np.float32][k], outs[np.float64][k], rtol=1e-5, what=f"{what} {k} t={t}")
    return worst


@pytest.mark.parametrize("name", ["ungm", "bot2", "bot1", "l63", "lg42", "lg42_tv", "msv"])
def test_gaussian_sum_filter_matches_reference(name):
    """inference.gaussian_sum_filter (inference.py:303-377)"""
    d = ref(f"gsf_{name}")
    m = omodel(name.replace("_tv", ""), d)
    M = int(d["M"])
    _mixture_stepwise(d, lambda w, pm, pP, y, u, Q, q0, R, r0: ogsf.gsf_step(m, w, pm, pP, y, u, Q, q0, R, r0), M, f"gsf {name}")
    if name in ("ungm", "lg42", "lg42_tv", "msv", "l63"):       # whole trajectories where fp32 rounding is not amplified
        T = d["emissions"].shape[0]
        g = ogsf.gaussian_sum_filter(m, d["emissions"], _inputs(d, T), d["means0"], d["P0"], d["Q"], d["q0"], d["R"], d["r0"])
        for k in FIELDS:
            close(g[k], d[k], rtol=2e-4, atol_scale=2e-5, what=f"gsf {name} trajectory {k}")


@pytest.mark.parametrize("name", ["ungm", "bot2", "l63", "lg42", "msv"])
def test_unscented_gaussian_sum_filter_matches_reference(name):
    """inference.unscented_gaussian_sum_filter (inference.py:379-456)"""
    d = ref(f"ugsf_{name}")
    m = omodel(name, d)
    up = ougsf.ParamsUKF(*[float(v) for v in d["uparams"]])
    _mixture_stepwise(d, lambda w, pm, pP, y, u, Q, q0, R, r0: ougsf.ugsf_step(m, w, pm, pP, y, u, Q, q0, R, r0, up),
                      int(d["M"]), f"ugsf {name}")


def test_kalman_step_matches_reference():
    """inference._kalman_step (inference.py:107-120)"""
    d = ref("kalman_step")
    for name in ("ungm", "bot2", "l63", "lg42"):
        m = omodel(name)
        res = {}
        for dt in (np.float32, np.float64):
            f = lambda a: np.asarray(a, dtype=dt)
            lls, pms, pPs = [], [], []
            for k in range(d[f"{name}_m"].shape[0]):
                ll, pm, pP = ogsf.kalman_step(m, f(d[f"{name}_m"][k])[None], f(d[f"{name}_P"][k])[None], f(d[f"{name}_Q"]),
                                              f(d[f"{name}_q0"]), f(d[f"{name}_R"]), f(d[f"{name}_r0"]), dt(d[f"{name}_u"][k]),
                                              f(d[f"{name}_y"][k]))
                lls.append(ll[0]); pms.append(pm[0]); pPs.append(pP[0])
            res[dt] = (np.asarray(lls), np.stack(pms), np.stack(pPs))
        for i, k in enumerate(("ll", "post_mean", "post_cov")):
            close_cond(d[f"{name}_{k}"], res[np.float32][i], res[np.float64][i], rtol=1e-5, what=f"kalman_step {name} {k}")


# ------------------------------------------------------------------------------------------ bootstrap particle filter
@pytest.mark.parametrize("name", ["ungm", "bot2", "lg3", "bot1", "msv"])
def test_bootstrap_particle_filter_matches_reference(name):
    """inference.bootstrap_particle_filter (inference.py:1302-1380) + utils._resample (utils.py:207-214), step by step from
    the reference's own carry.  Ancestors: exact on the reference's normalised weights; on the oracle's own weights (which
    differ from the reference's in the last bit) all but a handful of slots."""
    d = ref(f"bpf_{name}")
    m = omodel(name, d)
    N, T = int(d["N"]), d["emissions"].shape[0]
    u = _inputs(d, T)
    ess = float(d["ess_threshold"])
    W, X = d["weights"], d["particles"]
    n_res = 0
    for t in range(T):
        x = d["x0"] if t == 0 else X[:, t - 1]
        w = np.full(N, np.float32(1.0) / np.float32(N), np.float32) if t == 0 else W[:, t - 1]
        o32 = opf.pf_step(m, x, w, d["emissions"][t], u[t], d["noise"][t], d["res_u"][t], d["r0"], d["R"], ess, "multinomial")
        f64 = lambda a: np.asarray(a, np.float64)
        o64 = opf.pf_step(m, f64(x), f64(w), f64(d["emissions"][t]), float(u[t]), f64(d["noise"][t]), f64(d["res_u"][t]),
                          f64(d["r0"]), f64(d["R"]), ess, "multinomial")
        assert bool(o32["resampled"]) == bool(d["resampled"][t]), f"t={t}: ESS branch"
        if d["resampled"][t]:
            n_res += 1
            close_cond(d["norm_weights_on_resample"][t], o32["norm_weights"], o64["norm_weights"], rtol=1e-5, what=f"{name} w t={t}")
            assert np.array_equal(oscan.resample_ancestors(d["norm_weights_on_resample"][t], d["res_u"][t]), d["ancestors"][t])
            same = o32["ancestors"] == d["ancestors"][t]
            assert same.mean() >= 0.97, f"t={t}: {(~same).sum()} ancestors differ"
            close_cond(X[:, t][same], o32["particles"][same], o64["propagated"][d["ancestors"][t]][same], rtol=1e-5, what=f"{name} x t={t}")
            assert np.all(W[:, t] == np.float32(1.0) / np.float32(N))
        else:
            close_cond(W[:, t], o32["weights"], o64["weights"], rtol=1e-5, what=f"{name} w t={t}")
            close_cond(X[:, t], o32["particles"], o64["particles"], rtol=1e-5, what=f"{name} x t={t}")
    assert n_res > 0


# ------------------------------------------------------------------------------------------ augmented filter
@pytest.mark.parametrize("name", ["ungm", "bot2", "l63", "lg42", "tree_ungm", "tree_bot2"])
def test_speedy_agsf_matches_reference(name):
    """inference.speedy_augmented_gaussian_sum_filter (inference.py:621-812), step by step from the reference's carry;
    `tree_*`: outputs of inference.augmented_gaussian_sum_filter (:458-619, the GaussianComponent-tree version) -- the same
    recursion for the same draws, which is why ONE oracle function and one kernel serve both entry points."""
    d = ref(f"agsf_{name}")
    m = omodel(name.replace("tree_", ""))
    M, N1, N2 = (int(v) for v in d["num_components"])
    T = d["emissions"].shape[0]
    u = _inputs(d, T)
    for t in range(T):
        outs = {}
        for dt in (np.float32, np.float64):
            f = lambda a: np.asarray(a, dtype=dt)
            fm = f(d["means0"]) if t == 0 else f(d["means"][:, t - 1])
            fP = np.broadcast_to(f(d["P0"]), (M,) + d["P0"].shape).copy() if t == 0 else f(d["covariances"][:, t - 1])
            w = np.full(M, 1.0 / M, dtype=dt)
            outs[dt] = oagsf.agsf_step(m, fm, fP, w, f(d["emissions"][t]), dt(u[t]), f(d["Q"]), f(d["q0"]), f(d["R"]), f(d["r0"]),
                                       f(d["z_iid"][t]), f(d["s_iid"][t]), f(d["res_u"][t]), d["opt_args"])
        o32, o64 = outs[np.float32], outs[np.float64]
        close_cond(d["pre_weights"][t].reshape(-1), o32["pre_weights"], o64["pre_weights"], rtol=1e-5, what=f"{name} pre_weights t={t}")
        close_cond(d["updated_means"][t], o32["updated_means"], o64["updated_means"], rtol=1e-5, what=f"{name} updated_means t={t}")
        assert np.array_equal(oscan.resample_ancestors(d["pre_weights"][t].reshape(-1), d["res_u"][t]), d["res_idx"][t])
        idx = d["res_idx"][t]
        close_cond(d["means"][:, t], o32["updated_means"][idx], o64["updated_means"][idx], rtol=1e-5, what=f"{name} means t={t}")
        close_cond(d["covariances"][:, t], o32["updated_covs"][idx], o64["updated_covs"][idx], rtol=1e-5, what=f"{name} covs t={t}")
        assert np.all(d["weights"][:, t] == np.float32(1.0) / np.float32(M))


def test_optimal_resampling_matches_reference():
    """utils.optimal_resampling (utils.py:216-244)"""
    d = ref("optimal_resampling")
    k = 0
    while f"w{k}" in d.files:
        idx, w = oagsf.optimal_resampling(d[f"w{k}"], int(d[f"N{k}"]), d[f"u{k}"])
        assert np.array_equal(idx, d[f"idx{k}"]), k
        close(w, d[f"weights{k}"], rtol=1e-5, atol_scale=1e-6, what=f"optimal_resampling weights {k}")
        k += 1
    assert k >= 3


# ------------------------------------------------------------------------------------------ legacy NumPy classes (float64)
def _lmodel(name):
    return om.L63() if name == "l63" else om.UNGM(c=0.0)


@pytest.mark.parametrize("name", ["l63", "ungm"])
def test_legacy_bootstrap_pf_matches_reference(name):
    """particlefilt.BootstrapPF.run (particlefilt.py:26-52)"""
    d = ref("legacy_bpf")
    g = lambda k: d[f"{name}_{k}"]
    # the recorder drew SORTED uniforms u and took positions cdf[-1] * u; oracle.legacy takes positions cdf[-1] * (1 - v)
    parts = oleg.bootstrap_pf(_lmodel(name), g("ys"), g("Q"), g("R"), g("x0"), g("noise"), 1.0 - g("sorted_u"))
    np.testing.assert_allclose(parts, g("particles"), rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", ["l63", "ungm"])
def test_legacy_gausssumfilt_matches_reference(name):
    """gausssumfilt.GaussSumFilt.run (gausssumfilt.py:30-78): cov + F cov F^T, max shift, np.sum over all axes"""
    d = ref("legacy_gsf")
    g = lambda k: d[f"{name}_{k}"]
    out = oleg.gsf_legacy(_lmodel(name), g("ys"), g("Q"), g("R"), g("P0"), g("init_means"), True, True)
    for a, k in zip(out, ("means", "covs", "weights", "point")):
        np.testing.assert_allclose(a, g(k), rtol=1e-6, atol=1e-7, err_msg=k)      # the stand-in jacfwd is a 1e-9 finite difference


@pytest.mark.parametrize("name", ["l63", "ungm"])
@pytest.mark.parametrize("kind", ["EKF", "UKF", "UKFd", "MCF"])
def test_legacy_gaussfilt_matches_reference(name, kind):
    """gaussfilt.GaussFilt.run over EKF / UKF / MCF moment approximations (gaussfilt.py:88-198, :217-252)"""
    d = ref("legacy_gaussfilt")
    g = lambda k: d[f"{name}_{k}"]
    kw = dict(alpha=1.0, beta=2.0, kappa=0.5) if kind == "UKF" else {}
    ll, ms, Ps = oleg.gaussfilt_run(_lmodel(name), g("ys"), g("Q"), g("R"), g("m0"), g("P0"), kind.rstrip("d"),
                                    draws=g("MCF_draws") if kind == "MCF" else None, **kw)
    tol = 1e-4 if kind == "UKFd" else 1e-6       # alpha = 1e-3: weights of +-5e5 cancel, float64 keeps ~1e-5
    np.testing.assert_allclose(ms, g(f"{kind}_means"), rtol=tol, atol=tol, err_msg="means")
    np.testing.assert_allclose(Ps, g(f"{kind}_covs"), rtol=tol, atol=tol, err_msg="covs")
    np.testing.assert_allclose(ll, g(f"{kind}_ll"), rtol=tol, atol=tol, err_msg="ll")


def test_legacy_bank_matches_reference_and_records_its_defects():
    """gaussfilt.GaussSumFilt.run (gaussfilt.py:325-361) over EKF and UKF on the scalar model; on a 3-D model the class
    fails on its own [1,dx] reshapes, and MCLAF cannot run at all (sdp_opt arity) -- both recorded by the fixture."""
    d = ref("legacy_bank")
    g = lambda k: d[f"ungm_{k}"]
    for kind, kw in (("EKF", {}), ("UKF", dict(alpha=1.0, beta=2.0, kappa=0.5))):
        out = oleg.bank_run(_lmodel("ungm"), g("ys"), g("Q"), g("R"), g("P0"), g(f"{kind}_init_means"), kind, **kw)
        for a, k in zip(out, ("means", "covs", "weights", "point")):
            np.testing.assert_allclose(a, g(f"{kind}_{k}"), rtol=1e-6, atol=1e-7, err_msg=f"{kind} {k}")
    assert "l63_EKF_error" in d.files and "ungm_MCF_error" in d.files
    assert "sdp_opt" in str(ref("legacy_gaussfilt")["l63_mclaf_error"])


def test_utils_helpers_match_reference():
    """The package's own host-side helper faces (bayesianfiltering_b200/utils.py) against the reference's
    (tests/golden/ref_utils.npz, oracle/make_ref_fixtures.py gen_utils)."""
    from bayesianfiltering_b200 import utils as bu
    d = np.load(os.path.join(GOLD, "ref_utils.npz"))
    for k in range(3):
        np.testing.assert_allclose(bu.psd_solve(d[f"S{k}"], d[f"rhs{k}"]), d[f"psd_solve{k}"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(bu._get_sigma_points(d[f"m{k}"], d[f"P{k}"], d[f"lam{k}"]), d[f"sigma{k}"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(bu.split_to_sigma_points(d[f"m{k}"], d[f"P{k}"], float(d[f"lam{k}"])), d[f"legacy_sigma{k}"], rtol=1e-12, atol=1e-12)
        mean, cov = bu.collapse(d[f"cmeans{k}"], d[f"ccovs{k}"], d[f"cw{k}"])
        np.testing.assert_allclose(mean, d[f"collapse_mean{k}"], rtol=1e-12); np.testing.assert_allclose(cov, d[f"collapse_cov{k}"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(bu.gaussian_logpdf(d[f"y{k}"].astype(np.float64), d[f"my{k}"].astype(np.float64), d[f"S{k}"].astype(np.float64)),
                                   d[f"logpdf{k}"], rtol=1e-12)
        np.testing.assert_allclose(bu.mse(d[f"xe{k}"], d[f"xb{k}"]), d[f"mse{k}"], rtol=1e-12)
        np.testing.assert_allclose(bu.rmse(d[f"xe{k}"], d[f"xb{k}"]), d[f"rmse{k}"], rtol=1e-12)


def test_containers_match_reference():
    """containers.py:43-161 run unmodified (tests/golden/ref_containers.npz): conversions and the branching of a mixture,
    with the reference's own normal draws."""
    from bayesianfiltering_b200 import containers as bc
    d = np.load(os.path.join(GOLD, "ref_containers.npz"))
    for k in range(3):
        comps = bc._gaussian_sum_to_components(bc.GaussianSum(d[f"means{k}"], d[f"covs{k}"], d[f"w{k}"]))
        back = bc._components_to_gaussian_sum(comps)
        assert np.array_equal(np.stack(back.means), d[f"back_means{k}"]) and np.array_equal(np.stack(back.covariances), d[f"back_covs{k}"])
        assert np.array_equal(np.stack(back.weights), d[f"back_w{k}"])
        for tag, fn in (("t1", bc._branches_from_tree1), ("t2", bc._branches_from_tree2)):
            n = int(d[f"{tag}_n{k}"])
            assert n == (bc.num_prt1 if tag == "t1" else bc.num_prt2)
            tree = fn(comps, list(d[f"{tag}_delta{k}"]), [n] * len(comps), 0, normals=d[f"{tag}_z{k}"])
            assert len(tree) == len(comps) and all(len(br) == n for br in tree)
            np.testing.assert_allclose(np.stack([np.stack([c.mean for c in br]) for br in tree]), d[f"{tag}_means{k}"], rtol=2e-6, atol=2e-6)
            assert np.array_equal(np.stack([np.stack([c.covariance for c in br]) for br in tree]), d[f"{tag}_covs{k}"])
            np.testing.assert_allclose(np.stack([np.stack([c.weight for c in br]) for br in tree]), d[f"{tag}_w{k}"], rtol=1e-7)
