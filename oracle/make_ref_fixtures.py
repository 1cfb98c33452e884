#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz by EXECUTING THE UNMODIFIED REFERENCE.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_ref_fixtures            (in the build container: needs /root/reference)

The reference source is imported from /root/reference as it lies; jax / tensorflow_probability are replaced by the
NumPy stand-in of oracle/refshim.py (its header states exactly what is executed as written and what is restated).
Each fixture holds the inputs a filter consumed -- model constants, emissions, inputs and every random draw, read
off the shim's tape -- and the outputs the reference function returned.  The models are the lambdas of the
reference's own drivers, retyped here as test input with their file:line.

Functions executed (gaussfiltax/...):
  models.NonlinearSSM.sample                       models.py:240-289      -> ref_sample_*.npz
  inference.bootstrap_particle_filter              inference.py:1302-1380 -> ref_bpf_*.npz
  inference.gaussian_sum_filter                    inference.py:303-377   -> ref_gsf_*.npz
  inference.unscented_gaussian_sum_filter          inference.py:379-456   -> ref_ugsf_*.npz
  inference.speedy_augmented_gaussian_sum_filter   inference.py:621-812   -> ref_agsf_*.npz
  inference.augmented_gaussian_sum_filter          inference.py:458-619   -> ref_agsf_tree_*.npz (layout of ref_agsf_*)
  inference._kalman_step                           inference.py:107-120   -> ref_kalman_step.npz
  utils.optimal_resampling                         utils.py:216-244       -> ref_optimal_resampling.npz
  utils.psd_solve / _get_sigma_points / split_to_sigma_points / collapse / gaussian_logpdf / mse / rmse -> ref_utils.npz
  containers._branches_from_tree1/2, _gaussian_sum_to_components, _components_to_gaussian_sum  containers.py:43-161 -> ref_containers.npz
  particlefilt.BootstrapPF.run                     particlefilt.py:26-52  -> ref_legacy_bpf.npz
  gausssumfilt.GaussSumFilt.run                    gausssumfilt.py:30-78  -> ref_legacy_gsf.npz
  gaussfilt.{EKF,UKF,MCF,MCLAF} via GaussFilt.run  gaussfilt.py:88-308    -> ref_legacy_gaussfilt.npz
  gaussfilt.GaussSumFilt.run                       gaussfilt.py:325-361   -> ref_legacy_bank.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()
import jax.numpy as jnp  # noqa: E402  (the stand-in)
import jax.random as jr  # noqa: E402
from tensorflow_probability.substrates.jax.distributions import MultivariateNormalFullCovariance as MVN  # noqa: E402
import gaussfiltax.inference as gf  # noqa: E402  (the reference, unmodified)
import gaussfiltax.utils as gutils  # noqa: E402
from gaussfiltax.inference import ParamsUKF  # noqa: E402
from gaussfiltax.models import ParamsNLSSM, ParamsBPF, NonlinearSSM  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
f32 = lambda a: np.asarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------------ the drivers' models
def model_ungm():
    """Experiment_TSP_2023.ipynb cell 2 lines 30-34 (f3, g3, g3lp); inputs line 86: u_t = 8 cos(1.2 t)."""
    R = 1.0 * jnp.eye(1)
    r0 = jnp.zeros(1)
    f3 = lambda x, q, u: x / 2. + 25. * x / (1 + jnp.power(x, 2)) + u + q
    g3 = lambda x, r, u: 0.8 * x + r
    g3lp = lambda x, y, u: MVN(loc=g3(x, r0, u), covariance_matrix=R).log_prob(y)
    return dict(name="ungm", f=f3, h=g3, lp=g3lp, m0=jnp.zeros(1), P0=1.0 * jnp.eye(1), q0=jnp.zeros(1), Q=1.0 * jnp.eye(1),
                r0=r0, R=R, inputs=lambda T: 8.0 * np.cos(1.2 * np.arange(T)), dims=(1, 1, 1, 1))


def model_bot(dy=2):
    """BOT_Experiment_script.py:19-47 (fManBOT, gBOT2 / gBOT, gBOTlp), maneuver inputs :46."""
    dt = 0.5
    mu0 = 1.0 * jnp.array([-0.05, 0.001, 0.7, -0.05])
    Sigma0 = 1.0 * jnp.array([[0.1, 0.0, 0.0, 0.0], [0.0, 0.005, 0.0, 0.0], [0.0, 0.0, 0.1, 0.0], [0.0, 0.0, 0.0, 0.01]])
    Q = 1e-5 * jnp.eye(2)
    R = 25 * 1e-6 * jnp.eye(dy)
    r0 = jnp.zeros(dy)
    FCV = 1.0 * jnp.array([[1, dt, 0, 0], [0, 1, 0, 0], [0, 0, 1, dt], [0, 0, 0, 1]])
    acc = 0.5
    Omega = lambda x, acc: 0.1 * acc / jnp.sqrt(x[1] ** 2 + x[3] ** 2)
    FCT = lambda x, a: jnp.array([[1, jnp.sin(dt * Omega(x, a)) / Omega(x, a), 0, -(1 - jnp.cos(dt * Omega(x, a))) / Omega(x, a)],
                                  [0, jnp.cos(dt * Omega(x, a)), 0, -jnp.sin(dt * Omega(x, a))],
                                  [0, (1 - jnp.cos(dt * Omega(x, a))) / Omega(x, a), 1, jnp.sin(dt * Omega(x, a)) / Omega(x, a)],
                                  [0, jnp.sin(dt * Omega(x, a)), 0, jnp.cos(dt * Omega(x, a))]])
    G = jnp.array([[0.5, 0], [1, 0], [0, 0.5], [0, 1]])
    fManBOT = lambda x, q, u: (0.5 * (u - 1) * (u - 2) * FCV - u * (u - 2) * FCT(x, acc) + 0.5 * u * (u - 1) * FCT(x, -acc)) @ x + G @ q
    gBOT = lambda x, r, u: jnp.atleast_1d(jnp.arctan2(x[2], x[0])) + r
    gBOT2 = lambda x, r, u: jnp.array([jnp.arctan2(x[2], x[0]), jnp.sqrt(x[0] ** 2 + x[2] ** 2)]) + r
    g = gBOT2 if dy == 2 else gBOT
    glp = lambda x, y, u: MVN(loc=g(x, r0, u), covariance_matrix=R).log_prob(y)
    maneuver = lambda T: np.array([1] * int(2 * T / 5) + [0] * int(T / 5) + [2] * (T - int(2 * T / 5) - int(T / 5)))
    return dict(name=f"bot{dy}", f=fManBOT, h=g, lp=glp, m0=mu0, P0=Sigma0, q0=jnp.zeros(2), Q=Q, r0=r0, R=R,
                inputs=maneuver, dims=(4, 2, dy, dy))


def model_l63():
    """Experiment_TSP_2023.ipynb cell 2 lines 16-28 (lorentz_63 Euler step, f63; g1 = c x.x + r)."""
    dt, sigma, rho, beta, c = 0.01, 10.0, 28.0, 2.667, 0.001
    def lorentz_63(x):
        return jnp.array([sigma * (x[1] - x[0]), x[0] * (rho - x[2]) - x[1], x[0] * x[1] - beta * x[2]])
    f63 = lambda x, q, u: x + dt * lorentz_63(x) + q
    g1 = lambda x, r, u: jnp.atleast_1d(c * jnp.dot(x, x)) + r
    R = 0.1 * jnp.eye(1)
    r0 = jnp.zeros(1)
    lp = lambda x, y, u: MVN(loc=g1(x, r0, u), covariance_matrix=R).log_prob(y)
    return dict(name="l63", f=f63, h=g1, lp=lp, m0=jnp.zeros(3), P0=jnp.eye(3), q0=jnp.zeros(3), Q=0.5 * jnp.eye(3), r0=r0, R=R,
                inputs=None, dims=(3, 3, 1, 1))


def model_lg(dx=3):
    """adaptive_experiment.py:59-64 (fLG = A x + q, gLG = H x + r)."""
    if dx == 3:
        A = f32(0.8 * np.eye(3)); H = f32(0.1 * np.eye(3) + 0.05)
        Q = jnp.eye(3); R = 0.1 * jnp.eye(3); m0 = jnp.zeros(3); P0 = jnp.eye(3)
    else:
        A = f32([[0.9, 0.1, 0, 0], [0, 0.9, 0.1, 0], [0, 0, 0.8, 0.1], [0.05, 0, 0, 0.7]])
        H = f32([[1.0, 0, 0.5, 0], [0, 1.0, 0, 0.5]])
        Q = jnp.array(0.2 * np.eye(4) + 0.05); R = jnp.array([[0.3, 0.1], [0.1, 0.2]]); m0 = jnp.ones(4); P0 = jnp.eye(4)
    dy = H.shape[0]
    fLG = lambda x, q, u: A @ x + q
    gLG = lambda x, r, u: H @ x + r
    r0 = jnp.zeros(dy)
    lp = lambda x, y, u: MVN(loc=gLG(x, r0, u), covariance_matrix=R).log_prob(y)
    return dict(name=f"lg{dx}{dy if dx == 4 else ''}", f=fLG, h=gLG, lp=lp, m0=m0, P0=P0, q0=jnp.zeros(A.shape[0]), Q=Q, r0=r0, R=R,
                inputs=None, dims=(A.shape[0], A.shape[0], dy, dy), A=A, H=H)


def model_msv():
    """adaptive_experiment.py:47-57: multivariate stochastic volatility with leverage; the emission noise enters
    MULTIPLICATIVELY (`glmsv`), and the particle filter's log-prob has a state-dependent covariance (`lmsvlp`)."""
    state_dim = 2
    Phi = 0.9 * jnp.eye(state_dim)
    beta = 0.5 * jnp.ones(state_dim)
    sigma = 2.0
    H0 = 0.5 * jnp.eye(state_dim)
    R = 0.25 * jnp.eye(state_dim)
    r0 = jnp.zeros(state_dim)
    fmsv = lambda x, q, u: Phi @ x + q
    glmsv = lambda x, r, u: u * beta * jnp.exp(x / sigma) * r + (1 - u) * (H0 @ x + r)
    def lmsvlp(x, y, u):
        M = u * jnp.diag(beta * jnp.exp(x / sigma)) + (1 - u) * jnp.eye(state_dim)
        return MVN(loc=glmsv(x, r0, u), covariance_matrix=M @ R @ M.T).log_prob(y)
    return dict(name="msv", f=fmsv, h=glmsv, lp=lmsvlp, m0=jnp.zeros(2), P0=jnp.eye(2), q0=jnp.zeros(2), Q=0.3 * jnp.eye(2),
                r0=r0, R=R, inputs=lambda T: (np.arange(T) % 3 != 0).astype(np.float32), dims=(2, 2, 2, 2),
                theta=f32([0.9, 0.9, 0.5, 0.5, 2.0, 0.5, 0.5]))


def params_of(m, bpf=False, Q=None, R=None):
    kw = dict(initial_mean=m["m0"], initial_covariance=m["P0"], dynamics_function=m["f"], dynamics_noise_bias=m["q0"],
              dynamics_noise_covariance=m["Q"] if Q is None else Q, emission_function=m["h"], emission_noise_bias=m["r0"],
              emission_noise_covariance=m["R"] if R is None else R)
    return ParamsBPF(emission_distribution_log_prob=m["lp"], **kw) if bpf else ParamsNLSSM(**kw)


def consts(m, T):
    u = None if m["inputs"] is None else f32(m["inputs"](T))
    d = dict(m0=f32(m["m0"]), P0=f32(m["P0"]), q0=f32(m["q0"]), Q=f32(m["Q"]), r0=f32(m["r0"]), R=f32(m["R"]),
             dims=np.asarray(m["dims"], np.int32))
    if u is not None:
        d["inputs"] = u
    return d, u


def sample(m, T, seed):
    """NonlinearSSM.sample (models.py:240-289) -> states, emissions and the draws it consumed."""
    dx, dq, dy, dr = m["dims"]
    d, u = consts(m, T)
    inputs = jnp.zeros((T, 1)) if u is None else jnp.array(u)
    refshim.tape_clear()
    states, emissions = NonlinearSSM(dx, dq, dy, dr).sample(params_of(m), jr.PRNGKey(seed), T, inputs=inputs)
    draws = [p["x"] for p in refshim.tape_take("mvn_sample")]
    z0, rr0 = draws[0], draws[1]
    q = np.stack([np.zeros(dq, np.float32)] + draws[2::2])          # row 0 unused (models.py:271-275)
    r = np.stack([rr0] + draws[3::2])
    d.update(states=f32(states), emissions=f32(emissions), z0=f32(z0), q_noise=f32(q), r_noise=f32(r))
    return d


def save(name, d):
    path = os.path.join(OUT, f"ref_{name}.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in d.items()})
    print(f"{path}: {os.path.getsize(path) / 1024:.1f} KB")


def gen_sample():
    for m, T, seed in ((model_ungm(), 40, 1), (model_bot(2), 30, 2), (model_l63(), 25, 3), (model_lg(4), 20, 4), (model_msv(), 24, 5)):
        save(f"sample_{m['name']}", sample(m, T, seed))


def gen_bpf():
    for m, T, N, ess, seed in ((model_ungm(), 40, 256, 0.5, 11), (model_bot(2), 25, 192, 1.0, 12), (model_lg(3), 20, 128, 0.5, 13),
                               (model_bot(1), 20, 100, 0.7, 14), (model_msv(), 24, 160, 0.5, 15)):
        dx, dq, dy, dr = m["dims"]
        s = sample(m, T, 100 + seed)
        d, u = consts(m, T)
        inputs = None if u is None else jnp.array(u)
        refshim.tape_clear()
        out = gf.bootstrap_particle_filter(params_of(m, bpf=True), jnp.array(s["emissions"]), N, jr.PRNGKey(seed), inputs, ess)
        draws = [p["x"] for p in refshim.tape_take("mvn_sample")]
        x0 = np.stack(draws[:N])                                     # inference.py:1372-1373
        noise = np.stack(draws[N:]).reshape(T, N, dq)                # models.py:83, one per particle per step
        ch = refshim.tape_take("choice")                             # utils.py:210, only on resampling steps
        W = f32(out["weights"]); X = f32(out["particles"])
        resampled = np.zeros(T, np.int32); res_u = np.zeros((T, N), np.float32); res_p = np.zeros((T, N), np.float32)
        anc = np.tile(np.arange(N, dtype=np.int32), (T, 1))
        k = 0
        for t in range(T):
            if k < len(ch) and np.all(W[:, t] == np.float32(1.0) / np.float32(N)):
                resampled[t] = 1; res_u[t] = ch[k]["u"]; res_p[t] = ch[k]["p"]; anc[t] = ch[k]["idx"]; k += 1
        assert k == len(ch), (k, len(ch))
        d.update(emissions=s["emissions"], x0=f32(x0), noise=f32(noise), res_u=res_u, norm_weights_on_resample=res_p,
                 ancestors=anc, resampled=resampled, weights=W, particles=X, ess_threshold=np.float32(ess), N=np.int32(N))
        if "theta" in m:
            d["theta"] = m["theta"]
        save(f"bpf_{m['name']}", d)


def _gsf_like(fn, m, T, M, seed, tv=False, **kw):
    s = sample(m, T, 200 + seed)
    d, u = consts(m, T)
    inputs = None if u is None else jnp.array(u)
    Q = R = None
    if tv:     # `_get_params`: a covariance with a leading T axis is indexed by the step (inference.py:21)
        sc = (1.0 + 0.5 * np.sin(np.arange(T)))[:, None, None].astype(np.float32)
        Q = jnp.array(sc * f32(m["Q"])[None]); R = jnp.array(sc[::-1] * f32(m["R"])[None])
        d["Q"], d["R"] = f32(Q), f32(R)
    refshim.tape_clear()
    post = fn(params_of(m, Q=Q, R=R), jnp.array(s["emissions"]), inputs=inputs, **kw)
    means0 = refshim.tape_take("mvn_sample")[0]["x"]                 # MVN(m0,P0).sample(M, PRNGKey(0)), :367
    d.update(emissions=s["emissions"], means0=f32(means0), M=np.int32(M))
    return d, post


def gen_gsf():
    for m, T, M, seed, tv in ((model_ungm(), 40, 8, 21, False), (model_bot(2), 25, 6, 22, False), (model_l63(), 25, 5, 23, False),
                              (model_lg(4), 20, 4, 24, False), (model_lg(4), 16, 3, 25, True), (model_bot(1), 20, 5, 26, False),
                              (model_msv(), 24, 5, 27, False)):
        d, post = _gsf_like(lambda p, y, inputs, **kw: gf.gaussian_sum_filter(p, y, M, 1, inputs), m, T, M, seed, tv)
        for k in post._fields:
            d[k] = f32(getattr(post, k))
        if "theta" in m:
            d["theta"] = m["theta"]
        save(f"gsf_{m['name']}{'_tv' if tv else ''}", d)


def gen_ugsf():
    for m, T, M, seed, up in ((model_ungm(), 40, 8, 31, (1, 0, 0)), (model_bot(2), 25, 6, 32, (1, 0, 0)), (model_l63(), 25, 5, 33, (1, 0, 0)),
                              (model_lg(4), 20, 4, 34, (1.0, 2.0, 0.5)), (model_msv(), 24, 5, 35, (1, 0, 0))):
        uparams = ParamsUKF(*up)
        d, post = _gsf_like(lambda p, y, inputs, **kw: gf.unscented_gaussian_sum_filter(p, uparams, y, M, 1, inputs), m, T, M, seed)
        for k in post._fields:
            d[k] = f32(getattr(post, k))
        d["uparams"] = f32(up)
        if "theta" in m:
            d["theta"] = m["theta"]
        save(f"ugsf_{m['name']}", d)


def gen_agsf():
    for m, T, nc, seed, oa in ((model_ungm(), 30, [6, 2, 3], 41, (0.8, 0.1)), (model_bot(2), 20, [5, 2, 2], 42, (0.9, 0.9)),
                               (model_l63(), 20, [4, 3, 2], 43, (0.8, 0.5)), (model_lg(4), 16, [4, 2, 2], 44, (0.5, 0.5))):
        s = sample(m, T, 300 + seed)
        d, u = consts(m, T)
        inputs = None if u is None else jnp.array(u)
        refshim.tape_clear()
        post, aux = gf.speedy_augmented_gaussian_sum_filter(params_of(m), jnp.array(s["emissions"]), nc, rng_key=jr.PRNGKey(seed),
                                                            opt_args=oa, inputs=inputs)
        means0 = refshim.tape_take("mvn_sample")[0]["x"]
        nrm = refshim.tape_take("normal")          # per step: z-sample [M,dx,N1], s-sample [M*N1,dx,N2] (the same every step)
        ch = refshim.tape_take("choice")           # per step: jr.choice(PRNGKey(0), M*N1*N2, (M,), p=weights)
        d.update(emissions=s["emissions"], means0=f32(means0), num_components=np.asarray(nc, np.int32), opt_args=f32(oa),
                 z_iid=f32(np.stack([n["z"] for n in nrm[0::2]])), s_iid=f32(np.stack([n["z"] for n in nrm[1::2]])),
                 res_u=f32(np.stack([c["u"] for c in ch])), res_idx=np.stack([c["idx"] for c in ch]).astype(np.int32),
                 pre_weights=f32(aux["pre_weights"]), updated_means=f32(aux["updated_means"]),
                 weights=f32(post.weights), means=f32(post.means), covariances=f32(post.covariances))
        save(f"agsf_{m['name']}", d)


def gen_agsf_tree():
    """inference.augmented_gaussian_sum_filter (inference.py:458-619): the same recursion written over GaussianComponent
    trees (containers._branches_from_tree1/2; branch counts hard-coded to num_prt1 = num_prt2 = 2, containers.py:13-14).
    Saved in the layout of the speedy filter's fixtures: its `jr.multivariate_normal(key, mean, P - Delta, (2,))` draws
    z [2,dx] per component become z_iid[m] = z^T, so that ONE oracle / kernel serves both entry points."""
    for m, T, M, seed, oa in ((model_ungm(), 20, 6, 51, (0.8, 0.1)), (model_bot(2), 12, 5, 52, (0.9, 0.9))):
        dx = m["dims"][0]
        nc = [M, 2, 2]
        s = sample(m, T, 300 + seed)
        d, u = consts(m, T)
        inputs = None if u is None else jnp.array(u)
        refshim.tape_clear()
        post, aux = gf.augmented_gaussian_sum_filter(params_of(m), jnp.array(s["emissions"]), nc, rng_key=jr.PRNGKey(seed),
                                                     opt_args=oa, inputs=inputs)
        means0 = refshim.tape_take("mvn_sample")[0]["x"]
        mv = refshim.tape_take("multivariate_normal")     # per step: M draws [2,dx] (tree 1), then 2 M draws [2,dx] (tree 2)
        ch = refshim.tape_take("choice")
        assert len(mv) == T * 3 * M and len(ch) == T
        per = [mv[t * 3 * M:(t + 1) * 3 * M] for t in range(T)]
        z_iid = np.stack([np.stack([c["z"].T for c in st[:M]]) for st in per])             # [T, M, dx, 2]
        s_iid = np.stack([np.stack([c["z"].T for c in st[M:]]) for st in per])             # [T, 2 M, dx, 2]
        assert z_iid.shape == (T, M, dx, 2) and s_iid.shape == (T, 2 * M, dx, 2)
        d.update(emissions=s["emissions"], means0=f32(means0), num_components=np.asarray(nc, np.int32), opt_args=f32(oa),
                 z_iid=f32(z_iid), s_iid=f32(s_iid),
                 res_u=f32(np.stack([c["u"] for c in ch])), res_idx=np.stack([c["idx"] for c in ch]).astype(np.int32),
                 pre_weights=f32(aux["pre_weights"]), updated_means=f32(aux["updated_means"]),
                 weights=f32(post.weights), means=f32(post.means), covariances=f32(post.covariances))
        save(f"agsf_tree_{m['name']}", d)


def gen_kalman_step():
    """_kalman_step (inference.py:107-120): predict-then-update of one Gaussian, on the drivers' models."""
    from jax import jacfwd
    rows = {}
    rng = np.random.default_rng(5)
    for m in (model_ungm(), model_bot(2), model_l63(), model_lg(4)):
        dx, dq, dy, dr = m["dims"]
        f, h = m["f"], m["h"]
        Fx, Hx, Fq, Hr = jacfwd(f, 0), jacfwd(h, 0), jacfwd(f, 1), jacfwd(h, 1)
        ms, Ps, ys, us, lls, pms, pPs = [], [], [], [], [], [], []
        for k in range(6):
            mean = f32(m["m0"]) + f32(rng.normal(0, 0.3, dx))
            A = rng.normal(0, 0.2, (dx, dx))
            P = f32(np.asarray(m["P0"]) * 0.5 + A @ A.T * 0.1)
            u = np.float32(k % 3) if m["inputs"] is not None else np.float32(0.0)
            y = f32(h(f(mean, m["q0"], u), m["r0"], u)) + f32(rng.normal(0, 0.05, dy))
            ll, pm, pP = gf._kalman_step(jnp.array(mean), jnp.array(P), f, Fx, Fq, m["Q"], m["q0"], u, h, Hx, Hr, m["R"], m["r0"], jnp.array(y))
            ms.append(mean); Ps.append(P); ys.append(y); us.append(u); lls.append(ll); pms.append(pm); pPs.append(pP)
        d, _ = consts(m, 1)
        d.pop("inputs", None)
        rows.update({f"{m['name']}_{k}": v for k, v in d.items()})
        rows.update({f"{m['name']}_m": f32(ms), f"{m['name']}_P": f32(Ps), f"{m['name']}_y": f32(ys), f"{m['name']}_u": f32(us),
                     f"{m['name']}_ll": f32(lls), f"{m['name']}_post_mean": f32(pms), f"{m['name']}_post_cov": f32(pPs)})
    save("kalman_step", rows)


def gen_optimal_resampling():
    """utils.optimal_resampling (utils.py:216-244, Fearnhead & Clifford 2003) on a few weight vectors."""
    rng = np.random.default_rng(6)
    rows = {}
    for k, (M, N) in enumerate(((12, 5), (20, 8), (16, 16 - 1), (9, 3))):
        w = rng.random(M) ** 3
        w = f32(w / w.sum())
        refshim.tape_clear()
        idx, ww = gutils.optimal_resampling(jnp.array(w), N, jr.PRNGKey(70 + k))
        ch = refshim.tape_take("choice")[0]
        rows.update({f"w{k}": w, f"N{k}": np.int32(N), f"u{k}": ch["u"], f"idx{k}": np.asarray(idx, np.int32), f"weights{k}": f32(ww)})
    save("optimal_resampling", rows)


def gen_utils():
    """The small helpers of gaussfiltax/utils.py the callers use next to the filters: psd_solve :256-259, _get_sigma_points
    :247-254, split_to_sigma_points :52-65, collapse :10-18, gaussian_logpdf :75-79, mse / rmse :179-187."""
    rng = np.random.default_rng(9)
    rows = {}
    for k, (dx, dy) in enumerate(((1, 1), (3, 2), (4, 4))):
        A = rng.normal(0, 1, (dx, dx)); P = f32(A @ A.T + 0.5 * np.eye(dx))
        B = rng.normal(0, 1, (dy, dy)); S = f32(B @ B.T + 0.3 * np.eye(dy))
        rhs = f32(rng.normal(0, 1, (dy, dx)))
        m = f32(rng.normal(0, 1, dx))
        lam = np.float32(0.7 * k - 0.2)
        means = f32(rng.normal(0, 1, (5, dx)))
        covs = np.stack([P * (1 + 0.1 * j) for j in range(5)])
        wts = rng.random(5); wts = wts / wts.sum()
        mo, co = gutils.collapse(np.asarray(means, np.float64), np.asarray(covs, np.float64), wts)
        y = f32(rng.normal(0, 1, dy)); my = f32(rng.normal(0, 1, dy))
        xe, xb = rng.normal(0, 1, (7, dx)), rng.normal(0, 1, (7, dx))
        rows.update({f"P{k}": P, f"S{k}": S, f"rhs{k}": rhs, f"m{k}": m, f"lam{k}": lam,
                     f"psd_solve{k}": f32(gutils.psd_solve(jnp.array(S), jnp.array(rhs))),
                     f"sigma{k}": f32(gutils._get_sigma_points(jnp.array(m), jnp.array(P), lam)),
                     f"legacy_sigma{k}": np.asarray(gutils.split_to_sigma_points(np.asarray(m, np.float64), np.asarray(P, np.float64), float(lam)), np.float64),
                     f"cmeans{k}": means, f"ccovs{k}": covs, f"cw{k}": wts, f"collapse_mean{k}": mo, f"collapse_cov{k}": co,
                     f"y{k}": y, f"my{k}": my, f"logpdf{k}": np.float64(gutils.gaussian_logpdf(np.asarray(y, np.float64), np.asarray(my, np.float64), np.asarray(S, np.float64))),
                     f"xe{k}": xe, f"xb{k}": xb, f"mse{k}": np.float64(gutils.mse(xe, xb)), f"rmse{k}": np.float64(gutils.rmse(xe, xb))})
    save("utils", rows)


def gen_containers():
    """containers.py:43-161: _gaussian_sum_to_components / _components_to_gaussian_sum and the branching of a mixture,
    `_branches_from_tree1` / `_branches_from_tree2` (every component split into num_prt1 / num_prt2 = 2 Gaussian particles of
    covariance Delta whose means are drawn from N(mean, P - Delta)).  The normal draws come off the shim's tape."""
    import gaussfiltax.containers as gc
    rng = np.random.default_rng(10)
    rows = {}
    for k, dx in enumerate((1, 3, 4)):
        M = 3
        means = f32(rng.normal(0, 1, (M, dx)))
        covs = []
        for _ in range(M):
            A = rng.normal(0, 1, (dx, dx)); covs.append(f32(A @ A.T + 0.5 * np.eye(dx)))
        covs = np.stack(covs)
        w = rng.random(M); w = f32(w / w.sum())
        comps = gc._gaussian_sum_to_components(gc.GaussianSum(jnp.array(means), jnp.array(covs), jnp.array(w)))
        back = gc._components_to_gaussian_sum(comps)
        rows.update({f"means{k}": means, f"covs{k}": covs, f"w{k}": w, f"back_means{k}": f32(np.stack(back.means)),
                     f"back_covs{k}": f32(np.stack(back.covariances)), f"back_w{k}": f32(np.stack(back.weights))})
        for tag, fn, nb in (("t1", gc._branches_from_tree1, gc.num_prt1), ("t2", gc._branches_from_tree2, gc.num_prt2)):
            deltas = f32(0.3 * covs) if tag == "t1" else f32(0.6 * covs)
            refshim.tape_clear()
            tree = fn(comps, [jnp.array(d) for d in deltas], [nb] * M, jr.PRNGKey(90 + k))
            draws = refshim.tape_take("multivariate_normal")
            rows.update({f"{tag}_delta{k}": deltas, f"{tag}_z{k}": np.stack([d["z"] for d in draws]),
                         f"{tag}_means{k}": f32(np.stack([np.stack([c.mean for c in br]) for br in tree])),
                         f"{tag}_covs{k}": f32(np.stack([np.stack([c.covariance for c in br]) for br in tree])),
                         f"{tag}_w{k}": f32(np.stack([np.stack([c.weight for c in br]) for br in tree])), f"{tag}_n{k}": np.int32(nb)})
    save("containers", rows)


# ------------------------------------------------------------------------------------------ legacy NumPy classes
class _RecordingRandom:
    """Stands in for `numpy.random` inside the legacy modules (`from numpy import random`): every draw comes from a
    seeded generator and is recorded, so that the kernels can be given the same draws."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.mvn, self.multinomial_u = [], []

    def multivariate_normal(self, mean, cov, size=None):
        mean = np.asarray(mean, np.float64); cov = np.asarray(cov, np.float64)
        try:
            L = np.linalg.cholesky(cov)
        except np.linalg.LinAlgError:        # numpy's own sampler factorises by SVD and tolerates an indefinite cov
            lam, V = np.linalg.eigh((cov + cov.T) / 2)
            L = V * np.sqrt(np.maximum(lam, 0.0))
        shp = () if size is None else ((size,) if np.isscalar(size) else tuple(size))
        z = self.rng.standard_normal(shp + mean.shape)
        x = mean + z @ L.T
        self.mvn.append(x - mean)
        return x

    def multinomial(self, n, pvals):
        """counts of n categorical draws = ancestors of n SORTED uniform positions on the cdf of pvals"""
        u = np.sort(self.rng.random(n))
        self.multinomial_u.append(u)
        cdf = np.cumsum(pvals)
        idx = np.minimum(np.searchsorted(cdf, cdf[-1] * u, "left"), len(pvals) - 1)
        return np.bincount(idx, minlength=len(pvals))

    def normal(self, loc=0.0, scale=1.0, size=None):
        return self.rng.normal(loc, scale, size)

    def choice(self, *a, **k):
        return self.rng.choice(*a, **k)


def _legacy_models():
    """single-argument f(x), g(x) as the legacy drivers pass them (exp_lorentz63.py:37-41, exp_particlefilt.py)."""
    dt, sigma, rho, beta, c = 0.01, 10.0, 28.0, 2.667, 0.001
    f63 = lambda x: x + dt * jnp.array([sigma * (x[1] - x[0]), x[0] * (rho - x[2]) - x[1], x[0] * x[1] - beta * x[2]])
    g63 = lambda x: jnp.atleast_1d(c * jnp.dot(x, x))
    fu = lambda x: x / 2. + 25. * x / (1 + x ** 2)
    gu = lambda x: 0.8 * x
    return dict(l63=(f63, g63, 3, 1, np.zeros(3), np.eye(3), 0.5 * np.eye(3), 0.1 * np.eye(1)),
                ungm=(fu, gu, 1, 1, np.zeros(1), np.eye(1), np.eye(1), np.eye(1)))


def _legacy_data(f, g, dx, dy, m0, P0, Q, R, T, seed):
    rng = np.random.default_rng(seed)
    x = m0 + np.linalg.cholesky(P0) @ rng.standard_normal(dx)
    ys = []
    for t in range(T):
        x = np.asarray(f(x), np.float64) + np.linalg.cholesky(Q) @ rng.standard_normal(dx)
        ys.append(np.asarray(g(x), np.float64) + np.linalg.cholesky(R) @ rng.standard_normal(dy))
    return np.stack(ys)


def gen_legacy():
    """The legacy NumPy classes, float64.  Two adjustments of the ENVIRONMENT, none of the source: `numpy.random` as seen by
    each legacy module is a recording generator (the draws become fixture inputs), and `gaussfilt.py` gets the name `utils`
    it uses but never binds (`import gaussfiltax.utils` at gaussfilt.py:7 binds `gaussfiltax`; as shipped, GaussFilt.run
    raises NameError at :117).  MCLAF cannot run at all: gaussfilt.py:286 calls utils.sdp_opt with 8 arguments, utils.py:110
    takes 7 -- recorded in the fixture as `mclaf_error`."""
    import warnings
    import gaussfiltax.particlefilt as pfm
    import gaussfiltax.gausssumfilt as gsm
    import gaussfiltax.gaussfilt as gfm
    gfm.utils = gutils
    old = refshim._X64[0]
    refshim._X64[0] = True           # the legacy classes are float64 NumPy
    warnings.simplefilter("ignore")
    try:
        rows_pf, rows_gsf, rows_gf, rows_bank = {}, {}, {}, {}
        for name, (f, g, dx, dy, m0, P0, Q, R) in _legacy_models().items():
            T = 15
            ys = _legacy_data(f, g, dx, dy, m0, P0, Q, R, T, 7)
            ssm = gfm.SSM(dx, dy, np.zeros(dx), Q, np.zeros(dy), R, f=f, g=g)
            common = {f"{name}_ys": ys, f"{name}_m0": m0, f"{name}_P0": P0, f"{name}_Q": Q, f"{name}_R": R}
            # particlefilt.BootstrapPF.run
            N = 64
            rec = _RecordingRandom(81)
            pfm.random = rec
            parts = pfm.BootstrapPF(ssm, N).run(ys, m0, P0)
            rows_pf.update(common)
            rows_pf.update({f"{name}_x0": rec.mvn[0] + m0, f"{name}_noise": np.stack(rec.mvn[1:]).reshape(T, N, dx),
                            f"{name}_sorted_u": np.stack(rec.multinomial_u), f"{name}_particles": parts})
            # gausssumfilt.GaussSumFilt.run
            M = 5
            rec = _RecordingRandom(82)
            gsm.random = rec
            means, covs, weights, point = gsm.GaussSumFilt(ssm, M).run(ys, m0, P0)
            rows_gsf.update(common)
            rows_gsf.update({f"{name}_init_means": np.stack(rec.mvn) + m0, f"{name}_means": means, f"{name}_covs": covs,
                             f"{name}_weights": weights, f"{name}_point": point})
            # gaussfilt.GaussFilt.run over each moment approximation, and the bank of filters over each
            rows_gf.update(common)
            rows_bank.update(common)
            NP = 48
            makers = dict(EKF=lambda: gfm.EKF(ssm, order=1), UKF=lambda: gfm.UKF(ssm, alpha=1.0, beta=2.0, kappa=0.5),
                          UKFd=lambda: gfm.UKF(ssm), MCF=lambda: gfm.MCF(ssm, NP))
            for cls, mk in makers.items():
                rec = _RecordingRandom(83)
                gfm.random = rec
                ll, fm, fP = mk().run(ys, m0, P0)
                rows_gf.update({f"{name}_{cls}_ll": np.asarray(ll), f"{name}_{cls}_means": np.asarray(fm), f"{name}_{cls}_covs": np.asarray(fP)})
                if cls == "MCF":      # per step: NP draws ~ N(m,P) for the prediction, NP for the update, as z = L^-1 (x - m) is not kept: store x - m
                    rows_gf[f"{name}_MCF_draws"] = np.stack(rec.mvn).reshape(T, 2, NP, dx)
                if cls == "UKFd":
                    continue
                rec = _RecordingRandom(84)
                gfm.random = rec
                try:       # gaussfilt.py:349-351 hands moment_approx a [1,dx] mean: only shape-agnostic f, g survive that
                    means, covs, weights, point = gfm.GaussSumFilt(mk(), M).run(ys, m0, P0)
                except Exception as e:      # noqa: BLE001
                    rows_bank[f"{name}_{cls}_error"] = np.array(f"{type(e).__name__}: {e}")
                    continue
                rows_bank.update({f"{name}_{cls}_init_means": np.stack(rec.mvn[:M]) + m0, f"{name}_{cls}_means": means,
                                  f"{name}_{cls}_covs": covs, f"{name}_{cls}_weights": weights, f"{name}_{cls}_point": point})
                if cls == "MCF":      # after the M initial draws: per step, per component, pred then upd
                    rows_bank[f"{name}_MCF_draws"] = np.stack(rec.mvn[M:]).reshape(T, M, 2, NP, dx)
            try:
                gfm.random = _RecordingRandom(85)
                gfm.MCLAF(ssm, 8).run(ys, m0, P0)
                rows_gf[f"{name}_mclaf_error"] = np.array("")
            except Exception as e:      # noqa: BLE001
                rows_gf[f"{name}_mclaf_error"] = np.array(f"{type(e).__name__}: {e}")
        save("legacy_bpf", rows_pf)
        save("legacy_gsf", rows_gsf)
        save("legacy_gaussfilt", rows_gf)
        save("legacy_bank", rows_bank)
    finally:
        refshim._X64[0] = old


if __name__ == "__main__":
    which = sys.argv[1:] or ["sample", "bpf", "gsf", "ugsf", "agsf", "agsf_tree", "kalman_step", "optimal_resampling", "utils", "containers", "legacy"]
    os.makedirs(OUT, exist_ok=True)
    for w in which:
        globals()["gen_" + w]()
