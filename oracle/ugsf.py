"""Unscented Gaussian-sum filter oracle (SURVEY section 8(f) rank 2 -- the next row after the hot path).
TEST INFRASTRUCTURE ONLY: the checker of csrc/ugsf_kernels.cuh.  Pinned like the rest of oracle/: tests/test_ref_pin.py holds it
to `tests/golden/ref_ugsf_*.npz`, outputs of the unmodified reference function run on the NumPy stand-in (oracle/refshim.py).

Restates `gaussfiltax/inference.py`: `ParamsUKF` :41-49, `_ukf_predict_nonadditive` :146-174,
`_ukf_condition_on_nonadditive` :198-224, `unscented_gaussian_sum_filter` :379-456 (its `_step` :407-440 is the
Gaussian-sum `_step` with the two EKF moment steps replaced), and `gaussfiltax/utils.py:247-254` (`_get_sigma_points`).

Third-party piece: `jax.scipy.linalg.sqrtm` (utils.py:250, `jnp.real(scipy.linalg.sqrtm(P))` with `scipy` imported from
jax, utils.py:6): the Schur method for the principal square root.  For the symmetric positive semi-definite matrices it
is given here the principal root is V diag(sqrt(lambda)) V^T, which `sym_sqrtm` computes with `eigh`; the two agree to
rounding.  Quirks kept from the reference:
  * the sigma points are  m +- sqrt(n + lambda) * rows of L^T  (utils.py:251-252), n = dx + noise dim of the
    AUGMENTED state (state and noise stacked, block-diagonal covariance, inference.py:164-165 / :210-211);
  * the centre point enters the mean with weight lambda / (n + lambda) through a separate evaluation f(m, q0, u) and the
    covariance through the outer product of (f(m, q0, u) - mean) with weight lambda/(n+lambda) + 1 - alpha^2 + beta
    (:171-173, :217-218);
  * the update solves with `psd_solve` (jitter 1e-6 on every element, utils.py:258) and uses the un-jittered S in
    P - K S K^T and in the log-likelihood (:220-223);
  * lambda = alpha^2 (n + kappa) - n (:162, :208); with alpha = 1e-3 the weights are ~ +-1e6 and cancel catastrophically
    in fp32 -- the reference's experiment script runs `ParamsUKF(1, 0, 0)` (BOT_Experiment_script.py:108).
"""
from typing import NamedTuple

import numpy as np

from .gsf import psd_solve
from .models import mvn_log_prob


class ParamsUKF(NamedTuple):
    """inference.py:41-49."""
    alpha: float = 1e-3
    beta: float = 2.0
    kappa: float = 0.0


def sym_sqrtm(P):
    """Principal square root of a (batch of) symmetric PSD matrices: V diag(sqrt(max(lambda, 0))) V^T."""
    dt = P.dtype
    lam, V = np.linalg.eigh(P.astype(np.float64))
    root = (V * np.sqrt(np.clip(lam, 0.0, None))[..., None, :]) @ np.swapaxes(V, -1, -2)
    return root.astype(dt)


def get_sigma_points(m, P, lam):
    """utils.py:247-254 for a batch: m [M,n], P [M,n,n] -> [M,2n,n]."""
    dt = m.dtype
    n = m.shape[-1]
    L = sym_sqrtm(P)
    s = dt.type(np.sqrt(n + lam))
    Lt = np.swapaxes(L, -1, -2)
    plus = m[:, None, :] + s * Lt
    minus = m[:, None, :] - s * Lt
    return np.concatenate([plus, minus], axis=1).astype(dt)


def _augment(m, P, n0, N):
    """(m, n0) stacked and blockdiag(P, N) for every component (inference.py:164-165, :210-211)."""
    dt = m.dtype
    M, dx = m.shape
    dn = n0.shape[0]
    mA = np.concatenate([m, np.broadcast_to(n0.astype(dt), (M, dn))], axis=1)
    PA = np.zeros((M, dx + dn, dx + dn), dtype=dt)
    PA[:, :dx, :dx] = P
    PA[:, dx:, dx:] = N.astype(dt)
    return mA, PA


def ukf_predict_nonadditive(model, m, P, u, Q, up, q0):
    """vmapped `_ukf_predict_nonadditive` (inference.py:146-174): m [M,dx], P [M,dx,dx]."""
    dt = m.dtype
    M, dx = m.shape
    dq = Q.shape[0]
    n = dx + dq
    lam = up.alpha ** 2 * (n + up.kappa) - n                                        # :162
    mA, PA = _augment(m, P, q0, Q)
    sp = get_sigma_points(mA, PA, lam)                                              # [M,2n,n]
    new = model.f(sp[..., :dx].reshape(-1, dx), sp[..., dx:].reshape(-1, dq), u).reshape(M, 2 * n, dx).astype(dt)
    centre = model.f(m, np.broadcast_to(q0.astype(dt), (M, dq)), u).astype(dt)
    w_s = dt.type(1.0 / (2.0 * (lam + n)))
    w_c = dt.type(lam / (lam + n))
    mu = (np.sum(new, axis=1) * w_s + centre * w_c).astype(dt)                      # :171
    d = new - mu[:, None, :]
    dc = centre - mu
    Sig = (np.einsum('mij,mik->mjk', d, d) * w_s +
           dt.type(lam / (lam + n) + 1.0 - up.alpha ** 2 + up.beta) * np.einsum('mi,mj->mij', dc, dc)).astype(dt)   # :172-173
    return mu, Sig


def ukf_condition_on_nonadditive(model, m, P, R, u, y, up, r0):
    """vmapped `_ukf_condition_on_nonadditive` (inference.py:198-224)."""
    dt = m.dtype
    M, dx = m.shape
    dr = r0.shape[0]
    n = dx + dr
    lam = up.alpha ** 2 * (n + up.kappa) - n                                        # :208
    mA, PA = _augment(m, P, r0, R)
    sp = get_sigma_points(mA, PA, lam)
    dy = y.shape[0]
    new = model.h(sp[..., :dx].reshape(-1, dx), sp[..., dx:].reshape(-1, dr), u).reshape(M, 2 * n, dy).astype(dt)
    centre = model.h(m, np.broadcast_to(r0.astype(dt), (M, dr)), u).astype(dt)
    w_s = dt.type(1.0 / (2.0 * (lam + n)))
    w_c = dt.type(lam / (lam + n))
    mu = (np.sum(new, axis=1) * w_s + centre * w_c).astype(dt)                      # :216
    d = new - mu[:, None, :]
    dc = centre - mu
    S = (np.einsum('mij,mik->mjk', d, d) * w_s +
         dt.type(lam / (lam + n) + 1.0 - up.alpha ** 2 + up.beta) * np.einsum('mi,mj->mij', dc, dc)).astype(dt)     # :217-218
    C = (np.einsum('mij,mik->mjk', d, sp[..., :dx] - m[:, None, :]) * w_s).astype(dt)                             # :221  [M,dy,dx]
    K = np.swapaxes(psd_solve(S, C), 1, 2)                                          # :222
    Pp = (P - K @ S @ np.swapaxes(K, 1, 2)).astype(dt)                              # :223
    mp = (m + (K @ (y.astype(dt) - mu)[..., None])[..., 0]).astype(dt)              # :224
    ll = mvn_log_prob(mu, S, y.astype(dt))
    return ll, mp, Pp


def ugsf_step(model, w, pm, pP, y, u, Q, q0, R, r0, up):
    """One `_step` (inference.py:407-440): unscented update of the carried prediction, reweight, unscented predict."""
    dt = pm.dtype
    lls, fm, fP = ukf_condition_on_nonadditive(model, pm, pP, R, u, y, up, r0)
    mx = np.max(lls)
    wu = (np.exp(lls - mx).astype(dt) * w).astype(dt)                                # :424-426
    z = np.sum(wu, dtype=dt)
    wn = (wu / z).astype(dt)                                                         # :427
    npm, npP = ukf_predict_nonadditive(model, fm, fP, u, Q, up, q0)                  # :430
    return dict(weights=wn, means=fm, covariances=fP, predicted_means=npm, predicted_covariances=npP, lls=lls,
                loglik_inc=dt.type(mx + np.log(z)))


def unscented_gaussian_sum_filter(model, emissions, inputs, means0, P0, Q, q0, R, r0, up=ParamsUKF()):
    """Whole filter (inference.py:379-456), outputs in the reference's [M,T,...] order; the initial component means
    (reference: `MVN(m0,P0).sample(M, PRNGKey(0))`, :448) are supplied."""
    dt = means0.dtype
    T = emissions.shape[0]
    M = means0.shape[0]
    w = np.full(M, dt.type(1.0) / dt.type(M), dtype=dt)                              # :450
    pm = means0.copy()
    pP = np.broadcast_to(P0.astype(dt), (M,) + P0.shape).copy()                      # :449
    keys = ("weights", "means", "covariances", "predicted_means", "predicted_covariances")
    acc = {k: [] for k in keys}
    ll = np.zeros(T, dtype=dt)
    get = lambda x, t: x[t] if np.ndim(x) == 3 else x                                # `_get_params`, :411-414
    for t in range(T):
        u = inputs[t] if inputs is not None else 0.0
        out = ugsf_step(model, w, pm, pP, emissions[t], u, get(Q, t), q0, get(R, t), r0, up)
        w, pm, pP = out["weights"], out["predicted_means"], out["predicted_covariances"]
        for k in keys:
            acc[k].append(out[k])
        ll[t] = out["loglik_inc"]
    res = {k: np.stack(v, axis=1) for k, v in acc.items()}
    res["loglik_inc"] = ll
    res["marginal_loglik"] = np.sum(ll, dtype=np.float64)
    return res
