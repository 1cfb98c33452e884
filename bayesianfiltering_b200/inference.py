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
                              engine: Optional[ParticleFilterEngine] = None, shard_particles: bool = False):
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
                                          key_to_seed(key), device)
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
