"""Monte-Carlo trial driver: the `for i in range(Nsim)` loops of the reference's experiment scripts
(docs/experiments/BOT_Experiment_script.py:89-181, adaptive_experiment.py:105) as ONE batched call per GPU.

Independent trials are the unit that shards across the GPUs of a box with no data-path collective:
rank r owns the contiguous block `shard_range(num_trials, world, r)`, runs it through a trial-batched engine, and
only the per-trial summaries (point-estimate RMSE, marginal log-likelihood, ...) are gathered at the end.
One process per GPU (`torchrun`), `torch.distributed` for the gather (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(num_trials, world, rank):
    """Contiguous, balanced block of trials of `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(num_trials), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def run_trials(fn, num_trials, group=None):
    """Run `fn(start, stop) -> dict of arrays with a leading (stop - start) axis` on this rank's block of trials and
    return the dict for ALL trials, concatenated in trial order, on every rank.  Without an initialised process group
    this is a plain call `fn(0, num_trials)`."""
    dist = _dist()
    if dist is None:
        return {k: np.asarray(v) for k, v in fn(0, int(num_trials)).items()}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    start, stop = shard_range(num_trials, world, rank)
    mine = {k: np.asarray(v) for k, v in fn(start, stop).items()} if stop > start else {}
    for k, v in mine.items():
        if v.shape[0] != stop - start:
            raise ValueError(f"summary {k!r} has leading axis {v.shape[0]}, expected {stop - start} trials")
    parts = [None] * world
    dist.all_gather_object(parts, mine, group=group)
    keys = sorted({k for p in parts for k in p})
    return {k: np.concatenate([p[k] for p in parts if k in p], axis=0) for k in keys}


def rmse(x_est, x_true):
    """sqrt(sum((x_est - x_true)^2) / T) per trial: gaussfiltax/utils.py:179-187 over a leading trial axis [B,T,dx]."""
    d = np.asarray(x_est, np.float64) - np.asarray(x_true, np.float64)
    return np.sqrt(np.sum(d * d, axis=(1, 2)) / d.shape[1])


def bootstrap_ci(values, num_resamples=1000, alpha=0.05, seed=0):
    """Mean and percentile-bootstrap confidence interval of per-trial values (BOT_Experiment_script.py:184-199)."""
    v = np.asarray(values, np.float64)
    v = v[np.isfinite(v)]
    if v.size == 0:
        return float("nan"), (float("nan"), float("nan"))
    rng = np.random.default_rng(seed)
    means = rng.choice(v, size=(num_resamples, v.size), replace=True).mean(axis=1)
    return float(v.mean()), (float(np.quantile(means, alpha / 2)), float(np.quantile(means, 1 - alpha / 2)))


def particle_filter_trials(params, emissions, num_particles, true_states=None, key=0, inputs=None, ess_threshold=0.5,
                           resampler="multinomial", group=None, device=None):
    """B independent bootstrap particle filters (emissions [B,T,dy]) sharded over the ranks: returns per-trial
    'marginal_log_likelihood' [B], point estimates 'means' [B,T,dx] and, with `true_states` [B,T,dx], 'rmse' [B]."""
    from . import inference as gf
    from .random import key_to_seed
    emissions = np.asarray(emissions, np.float32)

    def block(start, stop):
        # ONE seed for every rank; the block's global trial index goes into the Philox counters (trial_offset), so trial b
        # draws the same numbers whatever the number of ranks
        out = gf.bootstrap_particle_filter(params, emissions[start:stop], num_particles, key_to_seed(key), inputs,
                                           ess_threshold, resampler=resampler, outputs=("means", "log_likelihood"),
                                           num_trials=stop - start, device=device, trial_offset=start)
        res = {"means": out["means"], "marginal_log_likelihood": out["marginal_log_likelihood"]}
        if true_states is not None:
            res["rmse"] = rmse(out["means"], np.asarray(true_states)[start:stop])
        return res

    return run_trials(block, emissions.shape[0], group)


def gaussian_sum_filter_trials(params, emissions, num_components, true_states=None, key=0, inputs=None, group=None, device=None):
    """B independent Gaussian-sum filters sharded over the ranks: 'marginal_log_likelihood' [B], mixture means
    'means' [B,T,dx] and, with `true_states`, 'rmse' [B]."""
    from . import inference as gf
    from .random import key_to_seed
    emissions = np.asarray(emissions, np.float32)

    def block(start, stop):
        # initial component means per GLOBAL trial (generator keyed by (key, trial)): independent of the number of ranks
        m0 = np.asarray(params.initial_mean, np.float64)
        L = np.linalg.cholesky(np.asarray(params.initial_covariance, np.float64))
        means0 = np.stack([m0 + np.random.default_rng([key_to_seed(key) & 0xFFFFFFFF, b]).standard_normal((int(num_components), m0.size)) @ L.T
                           for b in range(start, stop)]).astype(np.float32)
        _, ex = gf.gaussian_sum_filter(params, emissions[start:stop], num_components, 1, inputs, init_means=means0,
                                       outputs=("log_likelihood", "mixture_mean"), num_trials=stop - start, device=device,
                                       return_extras=True)
        res = {"means": ex["mixture_mean"], "marginal_log_likelihood": ex["marginal_log_likelihood"]}
        if true_states is not None:
            res["rmse"] = rmse(ex["mixture_mean"], np.asarray(true_states)[start:stop])
        return res

    return run_trials(block, emissions.shape[0], group)
