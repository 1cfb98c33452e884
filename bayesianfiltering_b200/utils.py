"""Host-side helpers with the names of gaussfiltax/utils.py that the hot path's callers use.

`_resample` (utils.py:207-214) and `psd_solve` (utils.py:256-259) are on the hot path and live inside the
kernels; the functions here are their stand-alone faces plus the NumPy metrics the drivers call.
"""
import numpy as np

from .random import key_to_seed


def gaussian_logpdf(y, m, S):
    """utils.py:75-79 (legacy log-likelihood, Cholesky based)."""
    D = m.shape[0]
    L = np.linalg.cholesky(S)
    x = np.linalg.solve(L, np.transpose(y - m))
    return -0.5 * D * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))) - 0.5 * np.sum(x ** 2)


def mse(x_est, x_base):
    """utils.py:179-182"""
    T = x_est.shape[0]
    return np.sum((np.asarray(x_est) - np.asarray(x_base)) ** 2) / T


def rmse(x_est, x_base):
    """utils.py:184-187"""
    return np.sqrt(mse(x_est, x_base))


def _resample(weights, particles, key, *, uniforms=None, resampler="multinomial"):
    """utils.py:207-214: multinomial resampling; returns (uniform weights, resampled particles, next key).

    Runs `bf_resample` on the GPU.  `uniforms` [N] replaces the draw jr.choice makes from keys[0]."""
    import torch
    from .engine import resample_ancestors
    from .random import split
    keys = split(key, 2)
    w = torch.as_tensor(np.asarray(weights, dtype=np.float32)).cuda()
    n = w.numel()
    if uniforms is None:
        rng = np.random.default_rng(key_to_seed(keys[0]))
        uniforms = rng.random(n if resampler == "multinomial" else 1, dtype=np.float32)
    u = torch.as_tensor(np.asarray(uniforms, dtype=np.float32)).cuda()
    idx = resample_ancestors(w, u, resampler).cpu().numpy()
    resampled = np.take(np.asarray(particles), idx, axis=0)
    return np.ones(n, dtype=np.float32) / n, resampled, keys[1]


def optimal_resampling(weights, N, key, *, uniforms=None):
    """utils.py:216-244 (Fearnhead & Clifford 2003): keep the largest weights exactly, resample the small ones with a common
    threshold weight; returns (indices [N], weights [N]).  Runs `bf_optimal_resampling` on the GPU.  `uniforms` [M] replaces the
    draw its `jr.choice(key, M, (M,), p=res_weights)` makes."""
    import torch
    from .engine import optimal_resampling_device
    w = torch.as_tensor(np.asarray(weights, dtype=np.float32)).cuda()
    if uniforms is None:
        uniforms = np.random.default_rng(key_to_seed(key)).random(w.numel(), dtype=np.float32)
    u = torch.as_tensor(np.asarray(uniforms, dtype=np.float32)).cuda()
    idx, wo = optimal_resampling_device(w, int(N), u)
    return idx.cpu().numpy(), wo.cpu().numpy()
