"""Filter entry points with the signatures of gaussfiltax/inference.py.

    bootstrap_particle_filter(params, emissions, num_particles, key, inputs, ess_threshold)   inference.py:1302-1309
    gaussian_sum_filter(params, emissions, num_components, num_iter, inputs)                   inference.py:303-309

Positional order, defaults, returned field names and axis order (particles/components first, time second,
`swap_axes_on_values`, inference.py:25-26) are the reference's.  Keyword-only extensions select the
resampler, supply the random inputs, batch independent Monte-Carlo trials and pick outputs.  The body of
each `lax.scan` is one call into libbfilt.so; untagged Python callables raise (zoo.resolve) -- there is no
CPU fallback.

NumPy (or anything array-like) in -> NumPy out, through the `*_run_host` C entry points (copies inside);
torch CUDA tensors in -> torch CUDA tensors out, enqueued on torch's current stream without synchronising.
"""
from typing import NamedTuple, Optional, Any

import numpy as np

from . import zoo
from .engine import ParticleFilterEngine, GaussianSumFilterEngine, PF_OUTPUTS, PF_FINALS, GSF_OUTPUTS
from .models import ParamsNLSSM, ParamsBPF  # noqa: F401  (re-exported like the reference does)
from .random import PRNGKey, key_to_seed


class PosteriorGaussianSumFiltered(NamedTuple):
    """Marginals of the Gaussian-sum filtering posterior (gaussfiltax/inference.py:29-39)."""
    weights: Optional[Any] = None                 # [num_components, ntime]
    means: Optional[Any] = None                   # [num_components, ntime, state_dim]
    covariances: Optional[Any] = None             # [num_components, ntime, state_dim, state_dim]
    predicted_means: Optional[Any] = None         # [num_components, ntime, state_dim]
    predicted_covariances: Optional[Any] = None   # [num_components, ntime, state_dim, state_dim]


class ParamsUKF(NamedTuple):
    """Parameters of the unscented transform (gaussfiltax/inference.py:41-49): spread alpha, prior-knowledge beta, kappa."""
    alpha: float = 1e-3
    beta: float = 2
    kappa: float = 0


# Helper faces of inference.py:21-26 (inside the filters these are index arithmetic of the kernels / of zoo.host_model_from_params)
def _get_params(x, dim, t):
    """inference.py:21: the t-th slice of a parameter that carries a leading time axis, else the parameter itself."""
    x = np.asarray(x)
    return x[t] if x.ndim == dim + 1 else x


def _process_input(x, y):
    """inference.py:23: zeros [T,1] when no inputs are given."""
    return np.zeros((y, 1), np.float32) if x is None else x


def _MVN_log_prob(mean, cov, y):
    """inference.py:24: log N(y | mean, cov) with TFP's Cholesky arithmetic in fp32 (host face; in-kernel: `emission_log_prob`
    of csrc/pf_kernels.cuh and the log-likelihood of `ekf_condition`)."""
    from .models import MVN
    return MVN(mean, cov).log_prob(np.atleast_1d(np.asarray(y, np.float32)))


def swap_axes_on_values(outputs, axis1=0, axis2=1):
    """inference.py:25-26: swap two axes of every array in a dict (views)."""
    return {k: (v.transpose(axis1, axis2) if _is_torch(v) else np.swapaxes(v, axis1, axis2)) for k, v in outputs.items()}


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _swap(a, batched):
    """time-major [B,T,K,...] -> the reference's [K,T,...] (or [B,K,T,...]); a view, not a copy."""
    if a.ndim < 3:
        return a if batched else a[0]
    if _is_torch(a):
        a = a.transpose(1, 2)
    else:
        a = np.swapaxes(a, 1, 2)
    return a if batched else a[0]


def _pf_rename(name):
    return {"log_likelihood": "log_likelihood_increments"}.get(name, name)


def bootstrap_particle_filter(params: ParamsBPF, emissions, num_particles, key=PRNGKey(0), inputs=None,
                              ess_threshold: float = 0.5, *, resampler: str = "multinomial", noise=None,
                              resample_uniforms=None, init_particles=None, init_weights=None,
                              outputs=("weights", "particles"), num_trials=None, path: str = "auto", device=None,
                              engine: Optional[ParticleFilterEngine] = None, shard_particles: bool = False, trial_offset: int = 0):
    r"""Bootstrap particle filter for the nonlinear state space model (inference.py:1302-1380).

    Args (reference): params, emissions [T,dy], num_particles, key, inputs [T] or None, ess_threshold.
    Keyword-only extensions:
        resampler          'multinomial' (the reference: jr.choice, utils.py:210) or 'systematic'.
        noise              [T,N,dq] coloured process noise q0 + L_Q z (else Philox(key)).
        resample_uniforms  multinomial [T,N], systematic [T] (else Philox(key)).
        init_particles     [N,dx] (else N(m0,P0) from Philox(key)); init_weights [N] (else 1/N).
        outputs            any of weights, particles, propagated, norm_weights, ancestors, means,
                           log_likelihood, ess, resampled, final_particles, final_weights.
        num_trials         B independent filters; every array above then takes a leading/embedded B axis
                           (emissions [B,T,dy] or shared [T,dy], noise [T,B,N,dq], uniforms [T,B,N] / [T,B]).
        path               'auto' | 'resident' (one CTA per trial, shared memory) | 'streamed' (HBM pipeline).
        trial_offset       global index of trial 0 of this call (Monte-Carlo runs split over ranks: the in-kernel Philox
                           counters use trial_offset + b, so the draws of a trial do not depend on the split).
        shard_particles    True under torchrun: ONE filter whose num_particles are split over the ranks' GPUs
                           (engine.ShardedParticleFilterEngine); per-particle arguments and results are then the
                           local shard, per-step scalars come from rank 0.
    Returns: dict; 'weights' [N,T], 'particles' [N,T,dx] as in the reference (views of time-major buffers),
        extras keyed by name ('log_likelihood' is returned as 'log_likelihood_increments' [T] plus the scalar
        'marginal_log_likelihood').
    """
    if engine is None:
        _, hm = zoo.host_model_from_params(params)
        B = 1 if num_trials is None else int(num_trials)
        if shard_particles:
            from .engine import ShardedParticleFilterEngine
            engine = ShardedParticleFilterEngine(hm, int(num_particles), float(ess_threshold), resampler, key_to_seed(key), device)
        else:
            engine = ParticleFilterEngine(hm, int(num_particles), B, float(ess_threshold), resampler, path,
                                          key_to_seed(key), device, trial_offset=trial_offset)
        own = True
    else:
        own = False
    batched = num_trials is not None
    run = engine.run if _is_torch(emissions) else engine.run_host
    raw = run(emissions, inputs=inputs, init_particles=init_particles, init_weights=init_weights, noise=noise,
              resample_u=resample_uniforms, outputs=tuple(outputs))
    out = {}
    for name, a in raw.items():
        if name in PF_FINALS:
            out[name] = a if batched else a[0]
        elif name in ("means", "log_likelihood", "ess", "resampled"):
            out[_pf_rename(name)] = a if batched else a[0]
        else:
            out[name] = _swap(a, batched)
    if "log_likelihood" in raw:
        ll = raw["log_likelihood"]
        tot = ll.double().sum(dim=1) if _is_torch(ll) else ll.astype(np.float64).sum(axis=1)
        out["marginal_log_likelihood"] = tot if batched else tot[0]
    if own and _is_torch(emissions):
        import torch
        torch.cuda.current_stream().synchronize()     # the engine (and its scratch) is about to be freed
    if own:
        engine.close()
    return out


def gaussian_sum_filter(params: ParamsNLSSM, emissions, num_components: int = 1, num_iter: int = 1, inputs=None, *,
                        init_means=None, init_covariances=None, init_weights=None, key=PRNGKey(0),
                        outputs=("weights", "means", "covariances", "predicted_means", "predicted_covariances"),
                        num_trials=None, device=None, engine: Optional[GaussianSumFilterEngine] = None,
                        return_extras: bool = False):
    r"""Gaussian-sum filter, a mixture of extended Kalman filters (inference.py:303-377).

    `num_iter` is accepted and ignored, as in the reference.  The initial component means are drawn from
    N(m0, P0) (reference: `MVN(m0,P0).sample(M, jr.PRNGKey(0))`, :367) with NumPy's generator seeded by `key`
    unless `init_means` [M,dx] is given; `init_covariances` defaults to P0 replicated (:368), weights to 1/M.
    Returns PosteriorGaussianSumFiltered with [M,T,...] fields ([B,M,T,...] when num_trials is given);
    with return_extras=True also a dict with 'log_likelihood_increments', 'marginal_log_likelihood',
    'mixture_mean' when those were requested in `outputs`.
    """
    m, hm, Q_t, R_t = zoo.host_model_from_params(params, time_varying=True)
    B = 1 if num_trials is None else int(num_trials)
    batched = num_trials is not None
    M = int(num_components)
    if init_means is None:
        rng = np.random.default_rng(key_to_seed(key))
        L = np.linalg.cholesky(np.asarray(params.initial_covariance, dtype=np.float64))
        z = rng.standard_normal((B, M, m.dx))
        init_means = (np.asarray(params.initial_mean, dtype=np.float64) + z @ L.T).astype(np.float32)
    own = engine is None
    if own:
        engine = GaussianSumFilterEngine(hm, M, B, device)
    run = engine.run if _is_torch(emissions) else engine.run_host
    raw = run(emissions, inputs=inputs, init_means=init_means, init_covs=init_covariances, init_weights=init_weights,
              outputs=tuple(outputs), Q_t=Q_t, R_t=R_t)          # covariances with a leading T axis (:21, :337-340)
    fields = {k: _swap(raw[k], batched) for k in PosteriorGaussianSumFiltered._fields if k in raw}
    post = PosteriorGaussianSumFiltered(**fields)
    extras = {}
    if "log_likelihood" in raw:
        ll = raw["log_likelihood"]
        extras["log_likelihood_increments"] = ll if batched else ll[0]
        tot = ll.double().sum(dim=1) if _is_torch(ll) else ll.astype(np.float64).sum(axis=1)
        extras["marginal_log_likelihood"] = tot if batched else tot[0]
    if "mixture_mean" in raw:
        extras["mixture_mean"] = raw["mixture_mean"] if batched else raw["mixture_mean"][0]
    if own:
        if _is_torch(emissions):
            import torch
            torch.cuda.current_stream().synchronize()
        engine.close()
    return (post, extras) if return_extras else post


def unscented_gaussian_sum_filter(params: ParamsNLSSM, uparams: ParamsUKF, emissions, num_components: int = 1, num_iter: int = 1,
                                  inputs=None, *, init_means=None, init_covariances=None, init_weights=None, key=PRNGKey(0),
                                  outputs=("weights", "means", "covariances", "predicted_means", "predicted_covariances"),
                                  num_trials=None, device=None, return_extras: bool = False):
    r"""Unscented Gaussian-sum filter (inference.py:379-456): every component is updated and predicted with the unscented
    transform of the noise-augmented state (`_ukf_condition_on_nonadditive` :198-224, `_ukf_predict_nonadditive` :146-174)
    instead of a linearisation.  Same arguments, defaults, returned fields and axis order as the reference; the keyword-only
    extensions are those of `gaussian_sum_filter`.  Note the reference's default ParamsUKF (alpha = 1e-3) makes the sigma-point
    weights +-5e5 and is unusable in float32 -- there as here; its drivers pass ParamsUKF(1, 0, 0)
    (BOT_Experiment_script.py:108)."""
    m, hm = zoo.host_model_from_params(params)
    B = 1 if num_trials is None else int(num_trials)
    batched = num_trials is not None
    M = int(num_components)
    if init_means is None:
        rng = np.random.default_rng(key_to_seed(key))
        L = np.linalg.cholesky(np.asarray(params.initial_covariance, dtype=np.float64))
        z = rng.standard_normal((B, M, m.dx))
        init_means = (np.asarray(params.initial_mean, dtype=np.float64) + z @ L.T).astype(np.float32)
    engine = GaussianSumFilterEngine(hm, M, B, device, unscented=(uparams.alpha, uparams.beta, uparams.kappa))
    run = engine.run if _is_torch(emissions) else engine.run_host
    raw = run(emissions, inputs=inputs, init_means=init_means, init_covs=init_covariances, init_weights=init_weights,
              outputs=tuple(outputs))
    post = PosteriorGaussianSumFiltered(**{k: _swap(raw[k], batched) for k in PosteriorGaussianSumFiltered._fields if k in raw})
    extras = {}
    if "log_likelihood" in raw:
        ll = raw["log_likelihood"]
        extras["log_likelihood_increments"] = ll if batched else ll[0]
        tot = ll.double().sum(dim=1) if _is_torch(ll) else ll.astype(np.float64).sum(axis=1)
        extras["marginal_log_likelihood"] = tot if batched else tot[0]
    if "mixture_mean" in raw:
        extras["mixture_mean"] = raw["mixture_mean"] if batched else raw["mixture_mean"][0]
    if _is_torch(emissions):
        import torch
        torch.cuda.current_stream().synchronize()
    engine.close()
    return (post, extras) if return_extras else post


def _kalman_step(m, P, f, F_x, F_q, Q, q0, u, h, H_x, H_r, R, r0, y, *, device=None):
    r"""One extended-Kalman predict + update of a single Gaussian (inference.py:107-120): returns
    (ll, posterior_mean, posterior_cov).  `f` and `h` must be the registry functions of one model (zoo); the Jacobian
    arguments F_x, F_q, H_x, H_r of the reference's signature are accepted and ignored -- the kernel has the model's analytic
    Jacobians.  On the GPU this is `bf_gsf_step` with one component in predict-first order."""
    from . import _lib
    mf = getattr(f, "__bf_model__", None)
    if mf is None or getattr(h, "__bf_model__", None) is not mf:
        raise TypeError("f and h must be the dynamics / emission functions of one bayesianfiltering_b200.zoo model")
    dx = mf.dx
    hm = mf.host_model(q0=q0, Q=Q, r0=r0, R=R, m0=np.zeros(dx, np.float32), P0=np.eye(dx, dtype=np.float32))
    eng = GaussianSumFilterEngine(hm, 1, 1, device, flags=_lib.GSF_PREDICT_FIRST)
    out = eng.run_host(np.asarray(y, np.float32).reshape(1, -1), inputs=np.asarray([u], np.float32),
                       init_means=np.asarray(m, np.float32).reshape(1, 1, dx), init_covs=np.asarray(P, np.float32).reshape(1, 1, dx, dx),
                       outputs=("means", "covariances", "log_likelihood"))
    eng.close()
    return out["log_likelihood"][0, 0], out["means"][0, 0, 0], out["covariances"][0, 0, 0]


def speedy_augmented_gaussian_sum_filter(params: ParamsNLSSM, emissions, num_components, rng_key=PRNGKey(0), num_iter: int = 1,
                                         opt_args=(0.1, 0.1), inputs=None, *, init_means=None, init_covariances=None,
                                         z_iid=None, s_iid=None, resample_uniforms=None, num_trials=None, device=None):
    r"""Augmented Gaussian-sum filter (inference.py:621-812): `num_components = [M, N1, N2]`; per step every component is split
    into N1 Gaussian particles of covariance `opt_args[0] P`, predicted, split into N2 of covariance `opt_args[1] P^-`,
    updated, reweighted, and M survivors are drawn with `jr.choice(p=weights)` (the particle filter's resampler).

    Returns `(PosteriorGaussianSumFiltered(weights [M,T], means [M,T,dx], covariances [M,T,dx,dx]), aux_outputs)` like the
    reference; aux_outputs holds 'pre_weights' [T,M*N1*N2], 'updated_means' [T,M*N1*N2,dx] (its :775-776), plus
    'ancestors' [T,M] and 'log_likelihood_increments' [T].  The reference reuses ONE set of normal / uniform draws at every
    step (its keys are closure constants, :677, :718, :764); so does this: `z_iid` [M,dx,N1], `s_iid` [M*N1,dx,N2],
    `resample_uniforms` [M] (drawn from `rng_key` when not supplied; give them a leading T axis for per-step draws)."""
    from .engine import AugmentedGaussianSumFilterEngine
    m, hm = zoo.host_model_from_params(params)
    M, N1, N2 = (int(v) for v in num_components)
    B = 1 if num_trials is None else int(num_trials)
    batched = num_trials is not None
    rng = np.random.default_rng(key_to_seed(rng_key))
    if init_means is None:                                 # :795
        L = np.linalg.cholesky(np.asarray(params.initial_covariance, dtype=np.float64))
        init_means = (np.asarray(params.initial_mean, dtype=np.float64) + np.random.default_rng(0).standard_normal((B, M, m.dx)) @ L.T).astype(np.float32)
    if z_iid is None:
        z_iid = rng.standard_normal((M, m.dx, N1)).astype(np.float32)
    if s_iid is None:
        s_iid = rng.standard_normal((M * N1, m.dx, N2)).astype(np.float32)
    if resample_uniforms is None:
        resample_uniforms = np.random.default_rng(0).random(M, dtype=np.float32)      # jr.choice(jr.PRNGKey(0), ...), :764
    eng = AugmentedGaussianSumFilterEngine(hm, (M, N1, N2), B, opt_args, key_to_seed(rng_key), device)
    raw = eng.run_host(emissions, inputs=inputs, init_means=init_means, init_covs=init_covariances, z_iid=z_iid, s_iid=s_iid,
                       res_u=resample_uniforms,
                       outputs=("weights", "means", "covariances", "pre_weights", "updated_means", "ancestors", "log_likelihood"))
    eng.close()
    post = PosteriorGaussianSumFiltered(_swap(raw["weights"], batched), _swap(raw["means"], batched), _swap(raw["covariances"], batched))
    pick = (lambda a: a) if batched else (lambda a: a[0])
    aux = {"pre_weights": pick(raw["pre_weights"]), "updated_means": pick(raw["updated_means"]), "ancestors": pick(raw["ancestors"]),
           "log_likelihood_increments": pick(raw["log_likelihood"])}
    return post, aux


def augmented_gaussian_sum_filter(params: ParamsNLSSM, emissions, num_components, rng_key=PRNGKey(0), num_iter: int = 1,
                                  opt_args=(0.1, 0.1), inputs=None, **kw):
    r"""inference.py:458-619: the same recursion written over `containers.GaussianComponent` trees (`_branches_from_tree1/2`,
    containers.py:63-161).  Its branch means are `jr.multivariate_normal(key, mean, P - Delta)` = `mean + chol(P - Delta) z`,
    i.e. for the same standard-normal draws exactly the speedy filter -- which is what runs.  (Pinned: the unmodified tree
    version's outputs, tests/golden/ref_agsf_tree_*.npz, are reproduced by the speedy restatement step by step,
    tests/test_ref_pin.py.)"""
    return speedy_augmented_gaussian_sum_filter(params, emissions, num_components, rng_key, num_iter, opt_args, inputs, **kw)
