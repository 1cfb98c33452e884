"""Host-side helpers with the names of gaussfiltax/utils.py that the hot path's callers use.

`_resample` (utils.py:207-214) and `psd_solve` (utils.py:256-259) are on the hot path and live inside the
kernels; the functions here are their stand-alone faces plus the NumPy metrics and small-matrix helpers the drivers
call next to the filters (a handful of flops on host arrays; tests/golden/ref_utils.npz pins them to the reference).
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


def psd_solve(A, b):
    """utils.py:256-259: `solve(A + 1e-6, b)` -- the jitter goes on EVERY element of A, as in the reference; fp32.
    (Inside the filters this is `ekf_condition` of csrc/gsf_kernels.cuh.)"""
    A = np.asarray(A, dtype=np.float32) + np.float32(1e-6)
    return np.linalg.solve(A, np.asarray(b, dtype=np.float32)).astype(np.float32)


def _sym_sqrt(P):
    lam, V = np.linalg.eigh(np.asarray(P, dtype=np.float64))
    return (V * np.sqrt(np.clip(lam, 0.0, None))) @ V.T


def _get_sigma_points(m, P, ulambda):
    """utils.py:247-254: the 2 dx sigma points m +- sqrt(dx + lambda) * rows of sqrtm(P)^T (principal square root;
    no centre point -- the unscented filters add it themselves).  Returns [2 dx, dx]."""
    m = np.asarray(m, dtype=np.float32)
    dx = m.shape[0]
    step = (np.sqrt(np.float32(dx) + np.float32(ulambda)) * _sym_sqrt(P).T).astype(np.float32)
    return np.concatenate([m + step, m - step], axis=0)


def split_to_sigma_points(mean, cov, lamda):
    """utils.py:52-65 (legacy UKF): centre point, then mean +- sqrt(dx + lamda) * columns of chol(cov).  [2 dx + 1, dx]."""
    mean = np.asarray(mean, dtype=np.float64).reshape(-1)
    dx = mean.shape[0]
    root = np.sqrt(np.asarray(cov, dtype=np.float64).reshape(dx, dx)) if dx == 1 else np.linalg.cholesky(np.asarray(cov, dtype=np.float64))
    step = np.sqrt(dx + lamda) * root.T
    return np.concatenate([mean[None], mean + step, mean - step], axis=0)


def collapse(mean_mat, covariance_tens, weight_vec):
    """utils.py:10-18: moment-match a mixture by one Gaussian -> (mean [dx], cov [dx,dx])."""
    mu, P, w = (np.asarray(a, dtype=np.float64) for a in (mean_mat, covariance_tens, weight_vec))
    mean = w @ mu
    d = mu - mean
    return mean, np.einsum("m,mij->ij", w, P + d[:, :, None] * d[:, None, :])
