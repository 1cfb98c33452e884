"""Shared parity cases: the same model built twice -- as an oracle model (NumPy) and as registry params for
the CUDA path -- plus seeded synthetic data and random inputs.  Used by tests/, smoke() and bench.py."""
import numpy as np

from oracle import models as om


class Case:
    def __init__(self, name, omodel, zmodel_fn, m0, P0, Q, R, q0=None, r0=None, inputs_fn=None):
        self.name, self.omodel, self.zmodel_fn = name, omodel, zmodel_fn
        f = lambda a: np.asarray(a, dtype=np.float32)
        self.m0, self.P0, self.Q, self.R = f(m0), f(P0), f(Q), f(R)
        self.q0 = np.zeros(omodel.dq, np.float32) if q0 is None else f(q0)
        self.r0 = np.zeros(omodel.dr, np.float32) if r0 is None else f(r0)
        self.inputs_fn = inputs_fn

    def inputs(self, T):
        return None if self.inputs_fn is None else np.asarray(self.inputs_fn(T), dtype=np.float32)

    # ---- registry params (import lazily: the package is the product, the oracle is the checker)
    def params_nlssm(self):
        from bayesianfiltering_b200.models import ParamsNLSSM
        z = self.zmodel_fn()
        return ParamsNLSSM(self.m0, self.P0, z.dynamics_function, self.q0, self.Q, z.emission_function, self.r0, self.R)

    def params_bpf(self):
        from bayesianfiltering_b200.models import ParamsBPF
        z = self.zmodel_fn()
        return ParamsBPF(self.m0, self.P0, z.dynamics_function, self.q0, self.Q, z.emission_function, self.r0, self.R,
                         z.emission_log_prob)

    def host_model(self):
        from bayesianfiltering_b200 import zoo
        return zoo.host_model_from_params(self.params_nlssm())[1]

    # ---- synthetic data (float64 simulation, cast to float32) and random inputs
    def simulate(self, T, seed):
        rng = np.random.default_rng(seed)
        m = self.omodel
        LQ = np.linalg.cholesky(self.Q.astype(np.float64)); LR = np.linalg.cholesky(self.R.astype(np.float64))
        LP = np.linalg.cholesky(self.P0.astype(np.float64))
        u = self.inputs(T)
        x = (self.m0.astype(np.float64) + LP @ rng.standard_normal(m.dx))[None]
        xs, ys = [], []
        for t in range(T):
            ut = 0.0 if u is None else float(u[t])
            if t > 0:
                q = (self.q0 + LQ @ rng.standard_normal(m.dq))[None]
                x = m.f(x, q, ut)
            r = (self.r0 + LR @ rng.standard_normal(m.dr))[None]
            xs.append(x[0]); ys.append(m.h(x, r, ut)[0])
        return np.stack(xs).astype(np.float32), np.stack(ys).astype(np.float32)

    def pf_randomness(self, T, N, seed, resampler="multinomial"):
        rng = np.random.default_rng(seed)
        m = self.omodel
        LQ = np.linalg.cholesky(self.Q.astype(np.float64)); LP = np.linalg.cholesky(self.P0.astype(np.float64))
        x0 = (self.m0 + rng.standard_normal((N, m.dx)) @ LP.T).astype(np.float32)
        noise = (self.q0 + rng.standard_normal((T, N, m.dq)) @ LQ.T).astype(np.float32)
        if resampler == "multinomial":
            ru = rng.random((T, N), dtype=np.float32)
        else:
            ru = rng.random((T,), dtype=np.float32)
        return x0, noise, ru

    def gsf_means0(self, M, seed):
        rng = np.random.default_rng(seed)
        LP = np.linalg.cholesky(self.P0.astype(np.float64))
        return (self.m0 + rng.standard_normal((M, self.omodel.dx)) @ LP.T).astype(np.float32)


def _zoo():
    from bayesianfiltering_b200 import zoo
    return zoo


def case(name):
    if name == "ungm":      # UNGM with u_t = 8 cos(1.2 t)
        return Case("ungm", om.UNGM(), lambda: _zoo().ungm(), [0.0], [[1.0]], [[1.0]], [[1.0]],
                    inputs_fn=lambda T: 8.0 * np.cos(1.2 * np.arange(T)))
    if name == "ungm_q10":  # classic Q=10
        return Case("ungm_q10", om.UNGM(), lambda: _zoo().ungm(), [0.0], [[1.0]], [[10.0]], [[1.0]],
                    inputs_fn=lambda T: 8.0 * np.cos(1.2 * np.arange(T)))
    if name in ("bot2", "bot1"):   # BOT_Experiment_script.py:16-51 constants (R widened for dy=1 as test_inference.py)
        dy = 2 if name == "bot2" else 1
        maneuver = lambda T: np.array([1] * int(2 * T / 5) + [0] * int(T / 5) + [2] * (T - int(2 * T / 5) - int(T / 5)))
        return Case(name, om.BOT(dy=dy), lambda: _zoo().bot(emission_dim=dy),
                    [-0.05, 0.001, 0.7, -0.05], np.diag([0.1, 0.005, 0.1, 0.01]), 1e-5 * np.eye(2), 25e-6 * np.eye(dy),
                    inputs_fn=maneuver)
    if name == "l63":       # Experiment_TSP_2023.ipynb cell 2 (Q scaled down to keep fp32 tests well conditioned)
        return Case("l63", om.L63(), lambda: _zoo().lorenz63(), np.zeros(3), np.eye(3), 0.5 * np.eye(3), 0.1 * np.eye(1))
    if name == "lg3":       # adaptive_experiment.py:59-64 shape
        A, H = 0.8 * np.eye(3), 0.1 * np.eye(3, 3) + 0.05
        return Case("lg3", om.LG(A, H), lambda: _zoo().linear_gaussian(A, H), np.zeros(3), np.eye(3), np.eye(3), 0.1 * np.eye(3))
    if name == "lg8":       # config 5: d=8
        A = 0.8 * np.eye(8) + 0.05 * np.eye(8, k=1)
        H = 0.5 * np.eye(8)
        return Case("lg8", om.LG(A, H), lambda: _zoo().linear_gaussian(A, H), np.zeros(8), np.eye(8), np.eye(8), 0.1 * np.eye(8))
    if name == "lg42":
        A = np.array([[0.9, 0.1, 0, 0], [0, 0.9, 0.1, 0], [0, 0, 0.8, 0.1], [0.05, 0, 0, 0.7]])
        H = np.array([[1.0, 0, 0.5, 0], [0, 1.0, 0, 0.5]])
        Q = 0.2 * np.eye(4) + 0.05
        R = np.array([[0.3, 0.1], [0.1, 0.2]])
        return Case("lg42", om.LG(A, H), lambda: _zoo().linear_gaussian(A, H), np.ones(4), np.eye(4), Q, R)
    if name in ("l96", "l96s", "l96_6", "l96_10"):   # BASELINE configs[3]: Lorenz-96 d=40 (l96s: d=8 for quick checks), Q = 0.1 I
        d = {"l96": 40, "l96s": 8, "l96_6": 6, "l96_10": 10}[name]   # d = 6, 10: workspace sizes that are not multiples of 16 bytes
        return Case(name, om.L96(d), lambda: _zoo().lorenz96(d), 8.0 + 0.5 * np.sin(np.arange(d)), 0.5 * np.eye(d), 0.1 * np.eye(d),
                    0.5 * np.eye(d // 2))
    if name == "sin1d":     # Experiment A / ICASSP toy
        return Case("sin1d", om.SIN1D(1, 10.0, 1.0), lambda: _zoo().sin1d(1, 10.0, 1.0), [0.0], [[1.0]], [[1.0]], [[1.0]])
    if name == "sin3d":
        return Case("sin3d", om.SIN1D(3, 10.0, 0.5), lambda: _zoo().sin1d(3, 10.0, 0.5), np.zeros(3), np.eye(3), 0.5 * np.eye(3), [[0.5]])
    raise KeyError(name)


ALL_CASES = ["ungm", "bot2", "bot1", "l63", "lg3", "lg8", "lg42", "sin1d", "sin3d"]


def close(a, b, rtol=1e-5, atol_scale=1e-5, what=""):
    """|a-b| <= rtol*|b| + atol_scale*max|b|: relative tolerance 1e-5 at the scale of the field."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b)) if b.size else 0.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_scale * scale + 1e-30, err_msg=what)


def close_cond(a, o32, o64, rtol=1e-5, k=8.0, what=""):
    """The 1e-5 bar, made honest about conditioning.  `o32` is the oracle run in float32 (the reference's
    precision), `o64` the same restatement run in float64 on the same inputs.  Where the quantity is well
    conditioned |o32 - o64| is ~1e-7 of the field and this is the plain relative-1e-5 test; where the fp32
    recursion itself loses digits (log-likelihoods of size 1e3 under R = 25e-6, `P - K S K^T` cancelling)
    the kernel is allowed k times the error the fp32 oracle itself makes, element-wise and as a typical
    relative error of the field.  A kernel bug moves `a` away from BOTH oracles and still fails."""
    a = np.asarray(a, dtype=np.float64); o32 = np.asarray(o32, dtype=np.float64); o64 = np.asarray(o64, dtype=np.float64)
    scale = np.max(np.abs(o64)) if o64.size else 0.0
    d = np.abs(o32 - o64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(np.abs(o64) > 0, d / np.abs(o64), 0.0)
    typical = float(np.median(rel)) if rel.size else 0.0
    tol = rtol * scale + rtol * np.abs(o64) + k * d + k * typical * np.abs(o64) + 1e-30
    err = np.abs(a - o32)
    bad = ~(err <= tol)
    if np.any(bad):
        i = np.unravel_index(np.argmax(np.where(bad, err / tol, 0.0)), err.shape) if err.ndim else ()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} outside tolerance; worst at {i}: got {a[i]!r}, "
                             f"fp32 oracle {o32[i]!r}, fp64 oracle {o64[i]!r}, tol {tol[i] if tol.ndim else tol!r}")


def smoke_check():
    """One small PF run (resident path) + one small streamed PF run + one small GSF run vs the oracle."""
    from oracle import pf as opf, gsf as ogsf, scan as oscan
    from bayesianfiltering_b200.engine import ParticleFilterEngine, GaussianSumFilterEngine
    # The PF check is step-wise (every step re-derived by the oracle from the kernel's own previous state, ancestors
    # bit-exact): UNGM amplifies rounding (slope up to 25), so whole trajectories of two correct fp32 implementations
    # drift apart within a few steps.
    from tests.test_gpu_pf import _stepwise, ALL_OUT
    c = case("ungm")
    T, N, M = 12, 1000, 16
    _, ys = c.simulate(T, 0)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 1)
    for path in ("resident", "streamed"):
        eng = ParticleFilterEngine(c.host_model(), N, 1, 0.5, "multinomial", path, 0, 0)
        out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
        assert _stepwise(c, out, ys, u, x0, noise, ru, 0.5, "multinomial") > 0, path
        eng.close()
    # the large-N kernel in its throughput configuration (Philox noise, systematic resampling): finite, resamples every
    # step, and bit-reproducible
    cb = case("bot2")
    _, yb = cb.simulate(6, 3)
    ub = cb.inputs(6)
    runs = []
    for _ in range(2):
        eng = ParticleFilterEngine(cb.host_model(), 1 << 16, 1, 1.0, "systematic", "streamed", 7, 0)
        runs.append(eng.run_host(yb, inputs=ub, outputs=("means", "log_likelihood", "resampled")))
        eng.close()
    assert np.all(np.isfinite(runs[0]["log_likelihood"])) and np.all(runs[0]["resampled"] == 1)
    assert np.array_equal(runs[0]["log_likelihood"], runs[1]["log_likelihood"]) and np.array_equal(runs[0]["means"], runs[1]["means"])
    means0 = c.gsf_means0(M, 2)
    g = ogsf.gaussian_sum_filter(c.omodel, ys, u, means0, c.P0, c.Q, c.q0, c.R, c.r0)
    eng = GaussianSumFilterEngine(c.host_model(), M, 1, 0)
    out = eng.run_host(ys[:5], inputs=u[:5], init_means=means0, outputs=("weights", "means", "covariances"))
    close(out["means"][0][:, :, 0].T, g["means"][:, :5, 0], rtol=1e-4, atol_scale=1e-4, what="gsf means")
    close(out["weights"][0].T, g["weights"][:, :5], rtol=1e-4, atol_scale=1e-4, what="gsf weights")
    eng.close()
    # the unscented and the augmented Gaussian-sum filter: one step each against the oracle
    from oracle import ugsf as ougsf, agsf as oagsf
    from bayesianfiltering_b200.engine import AugmentedGaussianSumFilterEngine
    cb = case("bot2")
    _, yb = cb.simulate(3, 5)
    ub = cb.inputs(3)
    mb = cb.gsf_means0(4, 6)
    eng = GaussianSumFilterEngine(cb.host_model(), 4, 1, 0, unscented=(1.0, 0.0, 0.0))
    out = eng.run_host(yb[:1], inputs=ub[:1], init_means=mb[None], outputs=("weights", "means"))
    eng.close()
    f8 = lambda a: np.asarray(a, np.float64)
    o8 = ougsf.ugsf_step(cb.omodel, np.full(4, 0.25), f8(mb), np.broadcast_to(f8(cb.P0), (4, 4, 4)).copy(), f8(yb[0]), float(ub[0]),
                         f8(cb.Q), f8(cb.q0), f8(cb.R), f8(cb.r0), ougsf.ParamsUKF(1.0, 0.0, 0.0))
    close(out["means"][0, 0], o8["means"], rtol=1e-3, atol_scale=1e-4, what="ugsf means")
    rng = np.random.default_rng(7)
    z1, z2, ru = rng.standard_normal((4, 4, 2)).astype(np.float32), rng.standard_normal((8, 4, 2)).astype(np.float32), rng.random(4, dtype=np.float32)
    eng = AugmentedGaussianSumFilterEngine(cb.host_model(), (4, 2, 2), 1, (0.8, 0.5), 0, 0)
    out = eng.run_host(yb[:1], inputs=ub[:1], init_means=mb[None], z_iid=z1, s_iid=z2, res_u=ru, outputs=("means", "pre_weights", "ancestors"))
    eng.close()
    assert np.array_equal(out["ancestors"][0, 0], oscan.resample_ancestors(out["pre_weights"][0, 0], ru)), "agsf survivors not bit-exact"
    o8 = oagsf.agsf_step(cb.omodel, f8(mb), np.broadcast_to(f8(cb.P0), (4, 4, 4)).copy(), np.full(4, 0.25), f8(yb[0]), float(ub[0]), f8(cb.Q),
                         f8(cb.q0), f8(cb.R), f8(cb.r0), f8(z1), f8(z2), f8(ru), (0.8, 0.5))
    close(out["pre_weights"][0, 0], o8["pre_weights"], rtol=1e-3, atol_scale=1e-4, what="agsf leaf weights")
    # resampler bit-exactness on a random weight vector
    import torch
    from bayesianfiltering_b200.engine import resample_ancestors
    rng = np.random.default_rng(3)
    w = rng.random(5000).astype(np.float32); w /= w.sum(dtype=np.float32)
    uu = rng.random(5000, dtype=np.float32)
    a = resample_ancestors(torch.from_numpy(w).cuda(), torch.from_numpy(uu).cuda()).cpu().numpy()
    assert np.array_equal(a, oscan.resample_ancestors(w, uu)), "ancestors not bit-exact"
