"""GPU tier: bootstrap particle filter (bf_pf_* through the C ABI) against the oracle's restatement of
gaussfiltax/inference.py:1302-1380.

Contract (BASELINE.json north_star): ancestor indices BIT-EXACT given identical normalised weights and the
same uniforms; particles / weights / log-likelihood within relative 1e-5 (fp32) at the scale of the field.
The maps are nonlinear (UNGM's slope reaches 25), so one-ulp differences between CUDA's fused multiply-adds
and NumPy grow along a trajectory; the 1e-5 bar is therefore checked per step with the kernel's own previous
state as the oracle's input (`_stepwise`), and whole trajectories are compared at 1e-5 on the models that do
not amplify rounding (linear-Gaussian) and at a looser, stated tolerance elsewhere.
"""
import os

import numpy as np
import pytest

from oracle import gsf as ogsf, pf as opf, scan as oscan
from tests.parity_cases import case, close, close_cond

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ALL_OUT = ("weights", "particles", "propagated", "norm_weights", "ancestors", "means", "log_likelihood", "ess",
           "resampled", "final_particles", "final_weights")


def _engine(c, N, B=1, thr=0.5, res="multinomial", path="auto", seed=0):
    from bayesianfiltering_b200.engine import ParticleFilterEngine
    return ParticleFilterEngine(c.host_model(), N, B, thr, res, path, seed, 0)


def _stepwise(c, out, ys, u, x0, noise, ru, thr, res, b=0, B=1):
    """Every step of the kernel's trajectory re-derived by the oracle from the kernel's previous state."""
    T = ys.shape[0]
    N = x0.shape[0]
    x, w = x0, np.full(N, np.float32(1.0) / np.float32(N), np.float32)
    n_res = 0
    for t in range(T):
        ut = 0.0 if u is None else u[t]
        o = opf.pf_step(c.omodel, x, w, ys[t], ut, noise[t], ru[t], c.r0, c.R, thr, res)
        f8 = lambda v: np.asarray(v, dtype=np.float64)
        o8 = opf.pf_step(c.omodel, f8(x), f8(w), f8(ys[t]), ut, f8(noise[t]), ru[t], c.r0, c.R, thr, res)
        close(out["propagated"][b, t], o["propagated"], what=f"propagated t={t}")
        # normalised weights and the by-products, from the kernel's own previous state
        close_cond(out["norm_weights"][b, t], o["norm_weights"], o8["norm_weights"], what=f"norm_weights t={t}")
        # the increment is max + log(sum): its rounding is set by the size of the per-particle log-likelihoods
        ll_scale = max(1.0, abs(float(np.max(o["lls"]))), abs(float(o["loglik_inc"])))
        ll_tol = 1e-5 * ll_scale + 8 * abs(float(o["loglik_inc"]) - float(o8["loglik_inc"]))
        assert abs(float(out["log_likelihood"][b, t]) - float(o["loglik_inc"])) <= ll_tol, f"loglik t={t}"
        # ESS and the weighted mean are functions of the weights: check them against the kernel's OWN normalised
        # weights (tight), the weights themselves against the oracle above (conditioning-aware).  When two particles
        # share all the weight under R = 25e-6 the ESS moves by 1e-4 per ulp of atan2, in numpy as much as in CUDA.
        nw8 = f8(out["norm_weights"][b, t])
        close(out["ess"][b, t], 1.0 / np.sum(nw8 * nw8), rtol=2e-5, atol_scale=2e-5, what=f"ess t={t}")
        close(out["means"][b, t], np.sum(nw8[:, None] * f8(out["propagated"][b, t]), axis=0), rtol=2e-5, atol_scale=2e-5,
              what=f"mean t={t}")
        flag = bool(out["resampled"][b, t])
        margin = abs(float(o["ess"]) - thr * N) / (thr * N)
        if margin > 1e-3:
            assert flag == o["resampled"], f"resample decision differs at t={t}"
        if flag:
            n_res += 1
            # BIT-EXACT ancestors given the kernel's own normalised weights and the supplied uniforms
            anc = oscan.resample_ancestors(out["norm_weights"][b, t], ru[t], res)
            assert np.array_equal(out["ancestors"][b, t], anc), f"ancestors differ at t={t}"
            assert np.array_equal(out["particles"][b, t], out["propagated"][b, t][anc]), f"gather differs at t={t}"
            assert np.all(out["weights"][b, t] == np.float32(1.0) / np.float32(N))
        else:
            assert np.array_equal(out["ancestors"][b, t], np.arange(N))
            assert np.array_equal(out["particles"][b, t], out["propagated"][b, t])
            assert np.array_equal(out["weights"][b, t], out["norm_weights"][b, t])
        x, w = out["particles"][b, t], out["weights"][b, t]
    assert np.array_equal(out["final_particles"][b], x) and np.array_equal(out["final_weights"][b], w)
    return n_res


@pytest.mark.parametrize("path", ["resident", "streamed"])
@pytest.mark.parametrize("name,T,N,thr,res", [
    ("ungm", 40, 1000, 0.5, "multinomial"),        # BASELINE config 1 shape (N=1000), multinomial = the reference
    ("ungm", 30, 777, 0.5, "systematic"),          # north_star's systematic grid, ragged N
    ("bot2", 30, 1500, 1.0, "multinomial"),        # BOT_Experiment_script.py:151: resample every step
    ("bot1", 20, 600, 0.5, "systematic"),
    ("l63", 25, 900, 0.5, "multinomial"),
    ("lg3", 25, 640, 0.5, "multinomial"),
    ("lg8", 20, 512, 0.7, "multinomial"),          # config 5 model
    ("lg42", 20, 500, 0.5, "systematic"),
    ("sin1d", 25, 1000, 0.5, "multinomial"),
    ("sin3d", 20, 300, 0.5, "multinomial"),
])
def test_pf_stepwise_parity(name, T, N, thr, res, path):
    c = case(name)
    _, ys = c.simulate(T, 7)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 8, res)
    eng = _engine(c, N, 1, thr, res, path)
    out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
    n_res = _stepwise(c, out, ys, u, x0, noise, ru, thr, res)
    assert n_res > 0, "case never resampled: it does not exercise the scan/search/gather"
    eng.close()


@pytest.mark.parametrize("name,T,N,thr,res", [
    ("bot2", 8, 16384, 1.0, "systematic"),         # 8 tiles, 8 CTAs: the tau path (N = 2^k) across tiles and CTAs
    ("lg42", 10, 8192, 0.6, "systematic"),         # resampling and non-resampling steps alternate
    ("ungm", 10, 4096, 0.5, "multinomial"),        # unsorted positions: bisection of tau in global memory
    ("bot2", 6, 6000, 1.0, "systematic"),          # several tiles, N not a power of two: literal bisection of the cdf
    ("l63", 6, 32768, 1.0, "systematic"),
    ("bot2", 6, 32768, 1.0, "multinomial"),        # 16 tiles: two-level search (tile table in shared memory, then in-tile)
])
def test_pf_stepwise_parity_many_tiles(name, T, N, thr, res):
    """The same step-wise check on filters that span several tiles of the summation tree and several CTAs of the
    persistent kernel (the cases above fit one tile)."""
    c = case(name)
    _, ys = c.simulate(T, 7)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 8, res)
    eng = _engine(c, N, 1, thr, res, "streamed")
    out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
    n_res = _stepwise(c, out, ys, u, x0, noise, ru, thr, res)
    assert n_res > 0
    eng.close()


@pytest.mark.parametrize("path", ["resident", "streamed"])
@pytest.mark.parametrize("gold", ["pf_ungm", "pf_bot2", "pf_ungm_sys", "pf_lg3"])
def test_pf_golden_fixtures(gold, path):
    """The committed oracle trajectories (tests/golden/make_golden.py).  Linear-Gaussian: the whole trajectory at the
    north_star tolerance (no error growth).  Nonlinear models (UNGM's slope reaches 25, BOT under R = 25e-6) amplify a
    one-ulp difference into different resampling decisions within a few steps, so there every step is run from the
    fixture's own carry (teacher-forced) and held to 1e-5 -- the same check the reference-generated fixtures get
    (tests/test_gpu_ref.py)."""
    from tests.golden.make_golden import PF_CASES
    cname, T, N, thr, res = PF_CASES[gold]
    g = np.load(os.path.join(GOLD, gold + ".npz"))
    c = case(cname)
    if not cname.startswith("lg"):
        from tests.test_gpu_ref import pf_teacher_forced
        gg = {k: g[k] for k in g.files}
        gg.update(r0=c.r0, R=c.R)

        class G(dict):
            files = property(lambda self: list(self.keys()))
        assert pf_teacher_forced(c.host_model(), c.omodel, G(gg), path, thr, res) > 0
        return
    u = g["inputs"] if "inputs" in g.files else None
    eng = _engine(c, N, 1, thr, res, path)
    out = eng.run_host(g["emissions"], inputs=u, init_particles=g["x0"], noise=g["noise"], resample_u=g["res_u"],
                       outputs=ALL_OUT)
    eng.close()
    assert np.array_equal(out["resampled"][0], g["resampled"])
    tol = 1e-5
    close(out["log_likelihood"][0], g["loglik_inc"], rtol=tol, atol_scale=tol, what="loglik")
    close(out["ess"][0], g["ess"], rtol=10 * tol, atol_scale=10 * tol, what="ess")
    close(out["means"][0], g["means"], rtol=10 * tol, atol_scale=10 * tol, what="means")
    # reference axis order: weights [N,T], particles [N,T,dx]  (inference.py:1378)
    W = np.swapaxes(out["weights"][0], 0, 1)
    X = np.swapaxes(out["particles"][0], 0, 1)
    close(X, g["particles"], rtol=tol, atol_scale=tol, what="particles")
    close(W, g["weights"], rtol=10 * tol, atol_scale=10 * tol, what="weights")


def test_reference_entry_point_shapes_and_axis_order():
    """bootstrap_particle_filter(params, emissions, num_particles, key, inputs, ess_threshold) returns the
    reference's dict: weights [N,T], particles [N,T,dx] (inference.py:1359-1362, 1378)."""
    from bayesianfiltering_b200 import inference as gf
    from bayesianfiltering_b200.random import PRNGKey
    c = case("bot2")
    T, N = 12, 300
    _, ys = c.simulate(T, 1)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 2)
    post = gf.bootstrap_particle_filter(c.params_bpf(), ys, N, PRNGKey(0), u, 1.0, noise=noise, resample_uniforms=ru,
                                        init_particles=x0)
    assert set(post) == {"weights", "particles"}
    assert post["weights"].shape == (N, T) and post["particles"].shape == (N, T, 4)
    ref = opf.bootstrap_particle_filter(c.omodel, ys, u, x0, noise, ru, c.r0, c.R, 1.0)
    # point estimate as the driver computes it (BOT_Experiment_script.py:152)
    est = np.einsum("ijk,ij->jk", post["particles"], post["weights"])
    est_ref = np.einsum("ijk,ij->jk", ref["particles"], ref["weights"])
    close(est, est_ref, rtol=1e-3, atol_scale=1e-3)
    # Philox mode (no supplied randomness): same key -> same result, different key -> different
    a = gf.bootstrap_particle_filter(c.params_bpf(), ys, N, PRNGKey(3), u, 1.0)
    b = gf.bootstrap_particle_filter(c.params_bpf(), ys, N, PRNGKey(3), u, 1.0)
    d = gf.bootstrap_particle_filter(c.params_bpf(), ys, N, PRNGKey(4), u, 1.0)
    assert np.array_equal(a["particles"], b["particles"]) and not np.array_equal(a["particles"], d["particles"])
    np.testing.assert_allclose(a["weights"].sum(axis=0), 1.0, rtol=1e-5)


@pytest.mark.parametrize("path", ["resident", "streamed"])
def test_resident_and_streamed_agree_and_chunked_resume_is_exact(path):
    """The carry (weights, particles) is an explicit state: a run split in time equals the unsplit run bit for
    bit (SURVEY section 5 checkpoint/resume)."""
    c = case("lg42")
    T, N = 24, 700
    _, ys = c.simulate(T, 3)
    x0, noise, ru = c.pf_randomness(T, N, 4)
    outs = ("weights", "particles", "log_likelihood", "ancestors")
    e1 = _engine(c, N, 1, 0.6, "multinomial", path)
    full = e1.run_host(ys, init_particles=x0, noise=noise, resample_u=ru, outputs=outs)
    e1.close()
    e2 = _engine(c, N, 1, 0.6, "multinomial", path)
    k = 9
    a = e2.run_host(ys[:k], init_particles=x0, noise=noise[:k], resample_u=ru[:k], outputs=outs)
    b = e2.run_host(ys[k:], noise=noise[k:], resample_u=ru[k:], outputs=outs, t_offset=k, resume=True)
    e2.close()
    for name in outs:
        assert np.array_equal(np.concatenate([a[name], b[name]], axis=1), full[name]), name
    # The other kernel is the same filter, not the same bits: the resident kernel sums the weights per thread and CTA, the
    # streamed one over its canonical tree (so that a filter sharded over GPUs equals the one-GPU filter), and a last-bit
    # difference in a weight can move an ancestor.  Same estimates to fp32 round-off, and the same ancestors until such a flip.
    other = _engine(c, N, 1, 0.6, "multinomial", "streamed" if path == "resident" else "resident")
    o = other.run_host(ys, init_particles=x0, noise=noise, resample_u=ru, outputs=outs)
    other.close()
    first = int(np.argmax(np.any(o["ancestors"] != full["ancestors"], axis=2)[0])) if np.any(o["ancestors"] != full["ancestors"]) else T
    assert first >= 2, f"the two kernels' ancestors already differ at step {first}"
    assert np.mean(o["ancestors"][0, :first + 1] == full["ancestors"][0, :first + 1]) > 0.99
    close(o["log_likelihood"][:, :first + 1], full["log_likelihood"][:, :first + 1], rtol=1e-4, atol_scale=1e-5,
          what="log-likelihood of the two kernels up to the first ancestor flip")
    close(o["log_likelihood"].sum(), full["log_likelihood"].sum(), rtol=5e-2, atol_scale=0.0, what="marginal log-likelihood of the two kernels")
    close(o["particles"][:, :first], full["particles"][:, :first])


@pytest.mark.parametrize("path", ["resident", "streamed"])
def test_trial_batch_equals_loop_over_trials(path):
    """num_trials=B is the reference drivers' `for i in range(Nsim)` loop (BOT_Experiment_script.py:89)."""
    c = case("ungm")
    T, N, B = 15, 400, 3
    ys = np.stack([c.simulate(T, 10 + b)[1] for b in range(B)])
    u = c.inputs(T)
    rnd = [c.pf_randomness(T, N, 20 + b) for b in range(B)]
    x0 = np.stack([r[0] for r in rnd]); noise = np.stack([r[1] for r in rnd], axis=1); ru = np.stack([r[2] for r in rnd], axis=1)
    eng = _engine(c, N, B, 0.5, "multinomial", path)
    out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
    eng.close()
    for b in range(B):
        one = _engine(c, N, 1, 0.5, "multinomial", path)
        o = one.run_host(ys[b], inputs=u, init_particles=x0[b], noise=noise[:, b], resample_u=ru[:, b], outputs=ALL_OUT)
        one.close()
        for k in ALL_OUT:
            assert np.array_equal(o[k][0], out[k][b]), (k, b)
    _stepwise(c, out, ys[1], u, x0[1], noise[:, 1], ru[:, 1], 0.5, "multinomial", b=1, B=B)


def test_streamed_runs_are_bit_reproducible():
    """The persistent kernel hands chunks to warps in a fixed order and folds its partial statistics in a fixed order:
    repeated runs are identical bit for bit (also the cheapest race detector there is -- compute-sanitizer is closed
    on the GPU pool)."""
    c = case("bot2")
    T, N = 12, 1 << 16
    _, ys = c.simulate(T, 21)
    u = c.inputs(T)
    outs = ("means", "log_likelihood", "ess", "ancestors", "final_particles")
    ref = None
    for rep in range(4):
        eng = _engine(c, N, 1, 1.0, "systematic", "streamed", seed=77)
        out = eng.run_host(ys, inputs=u, outputs=outs)
        eng.close()
        if ref is None:
            ref = out
        else:
            for k in outs:
                assert np.array_equal(out[k], ref[k]), (k, rep)
    assert np.all(np.isfinite(ref["log_likelihood"]))


@pytest.mark.parametrize("name,N,thr", [("bot2", 1 << 16, 1.0), ("bot2", 1 << 15, 0.5), ("ungm", 1 << 14, 0.5),
                                        ("lg3", 1 << 13, 0.7), ("l63", 1 << 13, 1.0), ("lg8", 1 << 13, 1.0)])
def test_lean_instantiation_equals_general_kernel(name, N, thr, monkeypatch):
    """A call in the throughput configuration (Philox noise, systematic resampling, power-of-two N, summaries only)
    runs the LEAN instantiation of k_pf_fused, whose optional code paths are compiled out.  It must be the same
    arithmetic: bit-identical to the general instantiation (BF_PF_LEAN=0), which the step-wise tests tie to the
    oracle."""
    c = case(name)
    T = 14
    _, ys = c.simulate(T, 9)
    u = c.inputs(T)
    outs = ("means", "log_likelihood", "ess", "resampled", "final_particles", "final_weights")
    res = {}
    for lean in ("1", "0"):
        monkeypatch.setenv("BF_PF_LEAN", lean)
        eng = _engine(c, N, 1, thr, "systematic", "streamed", seed=5)
        res[lean] = eng.run_host(ys, inputs=u, outputs=outs)
        eng.close()
    for k in outs:
        assert np.array_equal(res["1"][k], res["0"][k]), k
    assert np.all(np.isfinite(res["1"]["log_likelihood"]))
    if thr == 1.0:
        assert np.all(res["1"]["resampled"] == 1)


@pytest.mark.parametrize("thr", [1.0, 0.5])
def test_lean_kernel_chunked_resume_is_exact(thr):
    """Throughput configuration (lean instantiation, Philox noise indexed by the absolute step): a run split in time
    equals the unsplit run bit for bit, also when the split falls between a resampling and the gather that applies it."""
    c = case("bot2")
    T, N, k = 13, 1 << 15, 5
    _, ys = c.simulate(T, 17)
    u = c.inputs(T)
    outs = ("means", "log_likelihood", "ess", "resampled", "final_particles", "final_weights")
    e1 = _engine(c, N, 1, thr, "systematic", "streamed", seed=9)
    full = e1.run_host(ys, inputs=u, outputs=outs)
    e1.close()
    e2 = _engine(c, N, 1, thr, "systematic", "streamed", seed=9)
    a = e2.run_host(ys[:k], inputs=u[:k], outputs=outs)
    b = e2.run_host(ys[k:], inputs=u[k:], outputs=outs, t_offset=k, resume=True)
    e2.close()
    for name in ("means", "log_likelihood", "ess", "resampled"):
        assert np.array_equal(np.concatenate([a[name], b[name]], axis=1), full[name]), name
    assert np.array_equal(b["final_particles"], full["final_particles"]) and np.array_equal(b["final_weights"], full["final_weights"])


def test_device_tensor_mode_matches_host_mode():
    import torch
    c = case("bot2")
    T, N = 10, 5000
    _, ys = c.simulate(T, 5)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 6)
    eng = _engine(c, N, 1, 1.0, "multinomial", "streamed")
    host = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=("particles", "ancestors"))
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dev = eng.run(cu(ys), inputs=cu(u), init_particles=cu(x0[None]), noise=cu(noise[:, None]), resample_u=cu(ru[:, None]),
                  outputs=("particles", "ancestors"))
    torch.cuda.synchronize()
    assert eng.launches() > 0
    for k in host:
        assert np.array_equal(dev[k].cpu().numpy(), host[k]), k
    eng.close()


@pytest.mark.parametrize("N", [1, 2, 33])
def test_tiny_particle_counts(N):
    c = case("ungm")
    T = 6
    _, ys = c.simulate(T, 1)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 2)
    for path in ("resident", "streamed"):
        eng = _engine(c, N, 1, 1.1, "multinomial", path)       # thr > 1: resample at every step
        out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
        _stepwise(c, out, ys, u, x0, noise, ru, 1.1, "multinomial")
        eng.close()


@pytest.mark.parametrize("N", [8, 16, 64, 128, 256, 2048, 4096])
@pytest.mark.parametrize("res", ["systematic", "multinomial"])
def test_small_power_of_two_counts_on_the_large_n_kernel(N, res):
    """Power-of-two N takes the tau path of the large-N kernel (counting search for systematic positions, two-level
    search for multinomial ones from two tiles up): partially filled chunks and tiles, one tile exactly, two tiles."""
    c = case("bot2")
    T = 6
    _, ys = c.simulate(T, 1)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 2, res)
    eng = _engine(c, N, 1, 1.0, res, "streamed")
    out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
    assert _stepwise(c, out, ys, u, x0, noise, ru, 1.0, res) == T
    eng.close()


def test_errors_are_loud():
    from bayesianfiltering_b200.engine import ParticleFilterEngine
    c = case("ungm")
    with pytest.raises(ValueError):
        ParticleFilterEngine(c.host_model(), 0, 1)
    with pytest.raises(ValueError):
        ParticleFilterEngine(c.host_model(), 1 << 20, 1, path="resident")       # does not fit shared memory
    eng = _engine(c, 64)
    with pytest.raises(ValueError):
        eng.run_host(np.zeros((4, 1), np.float32), resume=True)                 # no carry yet
    with pytest.raises(ValueError):
        eng.run_host(np.zeros((4, 1), np.float32), noise=np.zeros((3, 64, 1), np.float32))
    eng.close()


def test_philox_pf_loglik_converges_to_exact_kalman():
    """BASELINE config 5: BPF on a linear-Gaussian SSM (d=8) against the exact Kalman-filter log-likelihood.
    The PF estimate of log Z is consistent; its spread shrinks as N grows."""
    c = case("lg8")
    T = 30
    _, ys = c.simulate(T, 9)
    A, H = c.omodel.A, c.omodel.H
    _, _, ll = ogsf.kalman_filter_exact(A, H, c.Q, c.R, c.m0, c.P0, ys)
    exact = ll.sum()
    errs = {}
    for N, path in ((1 << 10, "resident"), (1 << 14, "streamed"), (1 << 18, "streamed")):
        vals = []
        for seed in range(3):
            eng = _engine(c, N, 1, 0.5, "systematic", path, seed=100 + seed)
            out = eng.run_host(ys, outputs=("log_likelihood",))
            vals.append(out["log_likelihood"][0].astype(np.float64).sum())
            eng.close()
        errs[N] = np.mean(np.abs(np.array(vals) - exact))
    # d=8, T=30: the estimator's spread is O(1) at N=2^18 and shrinks like N^-1/2
    assert errs[1 << 18] < 1.5, errs
    assert errs[1 << 18] < errs[1 << 14] < errs[1 << 10], errs


@pytest.mark.parametrize("name,T,N,thr", [
    ("bot2", 6, 4096, 1.0),            # two tiles
    ("bot2", 6, 32768, 1.0),
    ("ungm", 8, 65536, 0.5),           # resampling and non-resampling steps
    ("bot2", 4, 1 << 18, 1.0),         # the size from which the table is the default
])
def test_multinomial_cell_table(name, T, N, thr, monkeypatch):
    """Unsorted positions through the cell table (pf_fused.cuh begin_gather: lower bounds of tau at the Ng + 1 cell edges,
    built by the counting search; a position then searches only the tau of its cell): every ancestor equals the
    reference's searchsorted, step by step, with the caller's own uniforms (arbitrary floats: the in-cell search) --
    the Philox uniforms are multiples of 2^-24 and take the edge shortcut at N = 2^24 (the headline-size test)."""
    monkeypatch.setenv("BF_PF_CELL_TABLE", "1")
    c = case(name)
    _, ys = c.simulate(T, 7)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 8, "multinomial")
    ru = ru.copy()
    ru[:, :4] = np.array([0.0, 1.0 - 2.0 ** -24, 0.5, 2.0 ** -24], np.float32)       # w = 1 (the last edge), the first cell, an edge
    ru[:, 4:512] *= np.float32(0.999)                                                  # not multiples of 2^-24
    eng = _engine(c, N, 1, thr, "multinomial", "streamed")
    out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
    assert _stepwise(c, out, ys, u, x0, noise, ru, thr, "multinomial") > 0
    monkeypatch.setenv("BF_PF_CELL_TABLE", "0")
    eng2 = _engine(c, N, 1, thr, "multinomial", "streamed")
    out2 = eng2.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru, outputs=ALL_OUT)
    for k in ("ancestors", "log_likelihood", "particles"):
        np.testing.assert_array_equal(out[k], out2[k], err_msg=k)
    eng.close(); eng2.close()


def test_large_streamed_properties_2pow22():
    """Size-independent properties at a large N (the oracle is too slow for whole trajectories there):
    weights sum to one, ancestors in range and non-decreasing for systematic positions, every gathered particle is
    a propagated one, the log-likelihood of a batch equals the sum of its two halves' runs (chunked resume)."""
    c = case("bot2")
    T, N = 4, 1 << 22
    _, ys = c.simulate(T, 2)
    u = c.inputs(T)
    ru = np.random.default_rng(3).random(T, dtype=np.float32)
    eng = _engine(c, N, 1, 1.0, "systematic", "streamed", seed=5)
    out = eng.run_host(ys, inputs=u, resample_u=ru,
                       outputs=("norm_weights", "ancestors", "particles", "propagated", "log_likelihood", "ess"))
    eng.close()
    for t in range(T):
        w = out["norm_weights"][0, t].astype(np.float64)
        assert abs(w.sum() - 1.0) < 1e-4
        a = out["ancestors"][0, t]
        assert a.min() >= 0 and a.max() < N
        assert np.all(np.diff(a) >= -1)            # monotone up to last-bit cdf wiggles
        assert np.array_equal(out["particles"][0, t], out["propagated"][0, t][a])
        close(out["ess"][0, t], 1.0 / np.sum(w * w), rtol=1e-4, atol_scale=1e-4)
    # full-size steps against the oracle (vectorised NumPy, seconds): bit-exact ancestors
    for t in (0, T - 1):
        anc = oscan.resample_ancestors(out["norm_weights"][0, t], ru[t], "systematic")
        assert np.array_equal(out["ancestors"][0, t], anc)


@pytest.mark.parametrize("res", ["systematic", "multinomial"])
def test_headline_size_step_against_oracle_2pow24(res):
    """The bench configuration's own size, N = 2^24 (147 CTAs x 56 tiles per CTA, the staged tile-sum scan, tau windows at
    full size), general instantiation with every per-particle output: two steps on the bearings-only model, each
    re-derived by the vectorised oracle from the kernel's previous state -- ancestors bit-exact, propagated particles
    and normalised weights at 1e-5, log-likelihood increment, ESS and weighted mean."""
    c = case("bot2")
    T, N = 2, 1 << 24
    _, ys = c.simulate(T, 21)
    u = c.inputs(T)
    x0, noise, ru = c.pf_randomness(T, N, 22, res)
    eng = _engine(c, N, 1, 1.0, res, "streamed")
    out = eng.run_host(ys, inputs=u, init_particles=x0, noise=noise, resample_u=ru,
                       outputs=("propagated", "norm_weights", "ancestors", "log_likelihood", "ess", "means", "resampled", "final_particles"))
    eng.close()
    x, w = x0, np.full(N, np.float32(1.0) / np.float32(N), np.float32)
    f8 = lambda v: np.asarray(v, dtype=np.float64)
    for t in range(T):
        o = opf.pf_step(c.omodel, x, w, ys[t], u[t], noise[t], ru[t], c.r0, c.R, 1.0, res)
        o8 = opf.pf_step(c.omodel, f8(x), f8(w), f8(ys[t]), u[t], f8(noise[t]), ru[t], c.r0, c.R, 1.0, res)
        close(out["propagated"][0, t], o["propagated"], what=f"propagated t={t}")
        close_cond(out["norm_weights"][0, t], o["norm_weights"], o8["norm_weights"], what=f"norm_weights t={t}")
        ll_scale = max(1.0, abs(float(np.max(o["lls"]))), abs(float(o["loglik_inc"])))
        assert abs(float(out["log_likelihood"][0, t]) - float(o["loglik_inc"])) <= 1e-5 * ll_scale + 8 * abs(float(o["loglik_inc"]) - float(o8["loglik_inc"]))
        nw8 = f8(out["norm_weights"][0, t])
        close(out["ess"][0, t], 1.0 / np.sum(nw8 * nw8), rtol=1e-4, atol_scale=1e-4, what=f"ess t={t}")
        close(out["means"][0, t], np.sum(nw8[:, None] * f8(out["propagated"][0, t]), axis=0), rtol=1e-4, atol_scale=1e-4, what=f"mean t={t}")
        assert out["resampled"][0, t] == 1
        anc = oscan.resample_ancestors(out["norm_weights"][0, t], ru[t], res)
        assert np.array_equal(out["ancestors"][0, t], anc), f"ancestors differ at t={t}"
        x = out["propagated"][0, t][anc]
        del o, o8
    assert np.array_equal(out["final_particles"][0], x)


def test_lean_equals_general_at_headline_size_2pow24(monkeypatch):
    """lean == general, bit for bit, at the size the bench runs (the lean kernel emits summaries only)."""
    c = case("bot2")
    T, N = 5, 1 << 24
    _, ys = c.simulate(T, 23)
    u = c.inputs(T)
    outs = ("means", "log_likelihood", "ess", "resampled")
    res = {}
    for lean in ("1", "0"):
        monkeypatch.setenv("BF_PF_LEAN", lean)
        eng = _engine(c, N, 1, 1.0, "systematic", "streamed", seed=5)
        res[lean] = eng.run_host(ys, inputs=u, outputs=outs)
        eng.close()
    for k in outs:
        assert np.array_equal(res["1"][k], res["0"][k]), k
    assert np.all(np.isfinite(res["1"]["log_likelihood"])) and np.all(res["1"]["resampled"] == 1)


@pytest.mark.parametrize("path,N", [("resident", 512), ("streamed", 4096)])
def test_trial_offset_makes_split_monte_carlo_runs_identical(path, N):
    """Independent trials split over handles (ranks): with `trial_offset` the in-kernel Philox counters use the GLOBAL
    trial index, so trial b gives the same bits whether it ran in one handle of 4 trials or as trial 0 of a second handle."""
    from bayesianfiltering_b200.engine import ParticleFilterEngine
    c = case("ungm")
    T = 8
    ys = np.stack([c.simulate(T, 30 + b)[1] for b in range(4)])
    u = c.inputs(T)
    outs = ("means", "log_likelihood", "final_particles")
    mk = lambda B, off: ParticleFilterEngine(c.host_model(), N, B, 0.5, "systematic", path, 77, 0, trial_offset=off)
    e = mk(4, 0); full = e.run_host(ys, inputs=u, outputs=outs); e.close()
    e = mk(2, 0); a = e.run_host(ys[:2], inputs=u, outputs=outs); e.close()
    e = mk(2, 2); b = e.run_host(ys[2:], inputs=u, outputs=outs); e.close()
    for k in outs:
        assert np.array_equal(np.concatenate([a[k], b[k]], axis=0), full[k]), k
    assert not np.array_equal(full["means"][0], full["means"][2])
