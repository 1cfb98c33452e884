"""Gaussian-sum (EKF mixture) filter oracle.  TEST INFRASTRUCTURE ONLY.

Restates `gaussfiltax/inference.py`: `_predict` :51-70, `_condition_on` :72-105, `_kalman_step`
:107-120, `gaussian_sum_filter` :303-377, and `gaussfiltax/utils.py:256-259` (`psd_solve`).
The initial component means (reference: `MVN(m0,P0).sample(M, PRNGKey(0))`, :367) are supplied.
"""
import numpy as np

from .models import mvn_log_prob


def psd_solve(A, b):
    """utils.py:256-259: `A = A + 1e-6` adds the jitter to EVERY element, then LU solve."""
    dt = A.dtype
    return np.linalg.solve(A + dt.type(1e-6), b).astype(dt)


def predict(model, m, P, Q, q0, u):
    """vmapped `_predict` (inference.py:51-70): m [M,dx], P [M,dx,dx]."""
    dt = m.dtype
    M = m.shape[0]
    q0b = np.broadcast_to(q0.astype(dt), (M, model.dq))
    Fx = model.Fx(m, q0b, u)
    Fq = model.Fq(m, q0b, u)
    mu = model.f(m, q0b, u)
    Sig = (Fx @ P @ np.swapaxes(Fx, 1, 2) + Fq @ Q.astype(dt) @ np.swapaxes(Fq, 1, 2)).astype(dt)
    return mu, Sig


def condition_on(model, m, P, R, r0, u, y):
    """vmapped `_condition_on` (inference.py:72-105)."""
    dt = m.dtype
    M = m.shape[0]
    r0b = np.broadcast_to(r0.astype(dt), (M, model.dr))
    Hx = model.Hx(m, r0b, u)
    Hr = model.Hr(m, r0b, u)
    S = (Hr @ R.astype(dt) @ np.swapaxes(Hr, 1, 2) + Hx @ P @ np.swapaxes(Hx, 1, 2)).astype(dt)   # :99
    K = np.swapaxes(psd_solve(S, Hx @ P), 1, 2)                                                    # :100
    Pp = (P - K @ S @ np.swapaxes(K, 1, 2)).astype(dt)                                            # :101
    yhat = model.h(m, r0b, u)
    mp = (m + (K @ (y.astype(dt) - yhat)[..., None])[..., 0]).astype(dt)                           # :102
    ll = mvn_log_prob(yhat, S, y.astype(dt))                                                       # :103
    return ll, mp, Pp


def kalman_step(model, m, P, Q, q0, R, r0, u, y):
    """vmapped `_kalman_step` (inference.py:107-120): predict from (m, P), then condition the prediction on y."""
    pm, pP = predict(model, m, P, Q, q0, u)
    return condition_on(model, pm, pP, R, r0, u, y)


def gsf_step(model, w, pm, pP, y, u, Q, q0, R, r0):
    """One `_step` (inference.py:333-365): update on the carried prediction, reweight, then predict."""
    dt = pm.dtype
    lls, fm, fP = condition_on(model, pm, pP, R, r0, u, y)
    mx = np.max(lls)
    lw = np.exp(lls - mx).astype(dt)                   # :347-348
    wu = (lw * w).astype(dt)                           # :349
    z = np.sum(wu, dtype=dt)
    wn = (wu / z).astype(dt)                           # :350
    loglik_inc = dt.type(mx + np.log(z))               # new output
    npm, npP = predict(model, fm, fP, Q, q0, u)        # :353
    return dict(weights=wn, means=fm, covariances=fP, predicted_means=npm, predicted_covariances=npP,
                lls=lls, loglik_inc=loglik_inc)


def gaussian_sum_filter(model, emissions, inputs, means0, P0, Q, q0, R, r0):
    """Whole filter, outputs in the reference's [M,T,...] order (`swap_axes_on_values`, :372)."""
    dt = means0.dtype
    T = emissions.shape[0]
    M = means0.shape[0]
    w = np.full(M, dt.type(1.0) / dt.type(M), dtype=dt)           # :369
    pm = means0.copy()
    pP = np.broadcast_to(P0.astype(dt), (M,) + P0.shape).copy()   # :368
    keys = ("weights", "means", "covariances", "predicted_means", "predicted_covariances")
    acc = {k: [] for k in keys}
    ll = np.zeros(T, dtype=dt)
    # `_get_params = lambda x, dim, t: x[t] if x.ndim == dim + 1 else x` (inference.py:21), called with dim = 2 for all
    # four (:337-340): a covariance with a leading T axis is time-varying; a bias never is (1-D has ndim 1, 2-D has 2).
    get = lambda x, t: x[t] if np.ndim(x) == 3 else x
    for t in range(T):
        u = inputs[t] if inputs is not None else 0.0
        out = gsf_step(model, w, pm, pP, emissions[t], u, get(Q, t), q0, get(R, t), r0)
        w, pm, pP = out["weights"], out["predicted_means"], out["predicted_covariances"]
        for k in keys:
            acc[k].append(out[k])
        ll[t] = out["loglik_inc"]
    res = {k: np.stack(v, axis=1) for k, v in acc.items()}
    res["loglik_inc"] = ll
    res["marginal_loglik"] = np.sum(ll, dtype=np.float64)
    return res


def kalman_filter_exact(A, H, Q, R, m0, P0, emissions):
    """Closed-form Kalman filter in float64 for f=Ax+q, h=Hx+r, in the reference's convention:
    the prior N(m0,P0) is the prediction for y_0 (update first, then predict; inference.py:345,353,367-369).
    Q and R may carry a leading T axis (time-varying).  Returns filtered means/covs and per-step log-likelihood
    increments."""
    A, H, Q, R = (np.asarray(a, dtype=np.float64) for a in (A, H, Q, R))
    Qs, Rs = Q, R
    m = np.asarray(m0, dtype=np.float64); P = np.asarray(P0, dtype=np.float64)
    T = emissions.shape[0]
    dy = H.shape[0]
    ms, Ps, ll = [], [], np.zeros(T)
    for t in range(T):
        y = np.asarray(emissions[t], dtype=np.float64)
        Q = Qs[t] if Qs.ndim == 3 else Qs
        R = Rs[t] if Rs.ndim == 3 else Rs
        S = H @ P @ H.T + R
        K = np.linalg.solve(S, H @ P).T
        e = y - H @ m
        L = np.linalg.cholesky(S)
        z = np.linalg.solve(L, e)
        ll[t] = -0.5 * z @ z - np.sum(np.log(np.diag(L))) - 0.5 * dy * np.log(2 * np.pi)
        m = m + K @ e
        P = P - K @ S @ K.T
        ms.append(m); Ps.append(P)
        m = A @ m
        P = A @ P @ A.T + Q
    return np.stack(ms), np.stack(Ps), ll
