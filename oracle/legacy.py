"""Legacy NumPy classes of the reference, restated with the randomness supplied.  TEST INFRASTRUCTURE ONLY.

  bootstrap_pf        gaussfiltax/particlefilt.py:26-52   BootstrapPF.run
  gsf_legacy          gaussfiltax/gausssumfilt.py:30-78   GaussSumFilt.run        (cov + F cov F^T, sic)
  ekf_run             gaussfiltax/gaussfilt.py:88-130     GaussFilt.run with EKF.moment_approx (:217-252)
  gsf_bank            gaussfiltax/gaussfilt.py:325-361    GaussSumFilt.run over EKF.moment_approx
float64 like the reference; `model` is an oracle model, f(x) = model.f(x, 0, 0), g(x) = model.h(x, 0, 0).
"""
import numpy as np


def gaussian_logpdf(y, m, S):
    """gaussfiltax/utils.py:75-79"""
    D = m.shape[0]
    L = np.linalg.cholesky(S)
    x = np.linalg.solve(L, np.transpose(y - m))
    return -0.5 * D * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))) - 0.5 * np.sum(x ** 2)


def _f(model, x):
    x = np.atleast_2d(np.asarray(x, np.float64))
    return model.f(x, np.zeros((x.shape[0], model.dq)), 0.0)


def _g(model, x):
    x = np.atleast_2d(np.asarray(x, np.float64))
    return model.h(x, np.zeros((x.shape[0], model.dr)), 0.0)


def _Fx(model, x):
    x = np.atleast_2d(np.asarray(x, np.float64))
    return model.Fx(x, np.zeros((x.shape[0], model.dq)), 0.0)


def _Gx(model, x):
    x = np.atleast_2d(np.asarray(x, np.float64))
    return model.Hx(x, np.zeros((x.shape[0], model.dr)), 0.0)


def bootstrap_pf(model, ys, Q, R, x0, noise, sorted_positions_u):
    """particlefilt.py:26-52 with x0 [N,dx], noise [T,N,dx] ~ N(0,Q) and, per step, the uniforms whose multinomial
    positions `1-u` are ascending (multinomial counts == ancestors of sorted positions).  Returns particles [T+1,N,dx]."""
    T, N, dx = ys.shape[0], x0.shape[0], x0.shape[1]
    particles = np.zeros((T + 1, N, dx)); particles[T] = x0
    for t in range(T):
        new = _f(model, particles[t - 1]) + noise[t]
        ll = np.array([gaussian_logpdf(ys[t], g, R) for g in _g(model, new)])
        ll -= np.max(ll)
        w = np.exp(ll); w /= np.sum(w)
        cdf = np.cumsum(w)
        idx = np.minimum(np.searchsorted(cdf, cdf[-1] * (1.0 - sorted_positions_u[t]), "left"), N - 1)
        particles[t] = new[idx]
    return particles


def _ekf_step(model, mean, cov, y, Q, R, cov_plus_p):
    """predict then update of every component (gausssumfilt.py:55-70 / gaussfilt.py:347-355)."""
    F = _Fx(model, mean)
    pm = _f(model, mean)
    pP = (cov if cov_plus_p else Q) + F @ cov @ np.swapaxes(F, 1, 2)
    G = _Gx(model, pm)
    mu_y = _g(model, pm)
    Sy = R + G @ pP @ np.swapaxes(G, 1, 2)
    Cxy = pP @ np.swapaxes(G, 1, 2)
    K = Cxy @ np.linalg.inv(Sy)
    fm = pm + ((y - mu_y)[:, None, :] @ np.swapaxes(K, 1, 2))[:, 0]
    fP = pP - K @ Sy @ np.swapaxes(K, 1, 2)
    ll = np.array([gaussian_logpdf(np.reshape(y, [1, -1]), mu_y[m], Sy[m]) for m in range(mean.shape[0])])
    return fm, fP, ll


def gsf_legacy(model, ys, Q, R, P0, init_means, cov_plus_p=True, max_shift=True):
    """gausssumfilt.py:30-78 (cov_plus_p, max_shift) or gaussfilt.py:325-361 (Q in the prediction, no max shift)."""
    T, (M, dx) = ys.shape[0], init_means.shape
    means = np.zeros((T + 1, dx, M)); covs = np.zeros((T + 1, dx, dx, M)); weights = np.zeros((T + 1, M))
    means[T] = init_means.T; covs[T] = np.asarray(P0, np.float64)[:, :, None]; weights[T] = 1.0 / M
    point = np.zeros((T, dx))
    for t in range(T):
        fm, fP, ll = _ekf_step(model, means[t - 1].T, np.transpose(covs[t - 1], (2, 0, 1)), ys[t], Q, R, cov_plus_p)
        means[t] = fm.T; covs[t] = np.transpose(fP, (1, 2, 0))
        if max_shift:
            ll = ll - np.max(ll)
        weights[t] = np.exp(ll) * weights[t - 1]
        weights[t] /= np.sum(weights[t])
        point[t] = np.sum(np.multiply(means[t, :], weights[t]))
    return means, covs, weights, point


def ekf_run(model, ys, Q, R, m0, P0):
    """gaussfilt.py:88-130 with EKF.moment_approx: -> (ll [T], means [T,dx], covs [T,dx,dx])."""
    T = ys.shape[0]
    m = np.asarray(m0, np.float64).reshape(1, -1); P = np.asarray(P0, np.float64)[None]
    ll = np.zeros(T); ms, Ps = [], []
    for t in range(T):
        m, P, l = _ekf_step(model, m, P, ys[t], Q, R, False)
        ll[t] = l[0]; ms.append(m[0]); Ps.append(P[0])
    return ll, np.stack(ms), np.stack(Ps)


# ---------------------------------------------------------------------------------------------------------------
# gaussfilt.py moment approximations beyond the EKF (UKF :133-170, MCF :173-198) and the filters built on them.
def split_to_sigma_points(mean, cov, lamda):
    """gaussfiltax/utils.py:52-65: [mean, mean +- sqrt(dx+lamda) * columns of chol(cov)] (dx = 1: sqrt(cov))."""
    dx = mean.shape[0]
    pts = np.zeros((2 * dx + 1, dx))
    pts[0] = mean
    c = np.sqrt(dx + lamda)
    L = np.sqrt(cov) if dx == 1 else np.linalg.cholesky(cov)
    for i in range(dx):
        pts[i + 1] = mean + c * L.T[i]
        pts[dx + i + 1] = mean - c * L.T[i]
    return pts


def _step_fn(model, kw, Q, R):
    """GaussFilt.which_step (gaussfilt.py:132-143)"""
    if kw == 'pred':
        return (lambda X: _f(model, X)), np.asarray(Q, np.float64)
    return (lambda X: _g(model, X)), np.asarray(R, np.float64)


def ukf_moments(model, m, P, kw, Q, R, alpha=1e-3, beta=2.0, kappa=0.0):
    """UKF.moment_approx (gaussfilt.py:146-170) -> (mean_out [do], var_out [do,do], cov_out [dx,do])."""
    func, cov = _step_fn(model, kw, Q, R)
    dx = m.shape[0]
    lam = alpha ** 2 * (dx + kappa) - dx
    sp = split_to_sigma_points(m, P, lam)
    nsp = func(sp)
    wm0 = lam / (dx + lam)
    wi = 1.0 / (2 * (dx + lam))
    wc0 = wm0 + 1 - alpha ** 2 + beta
    mean_out = wm0 * nsp[0] + wi * np.sum(nsp[1:], axis=0)
    var_out = cov + wc0 * np.outer(nsp[0] - mean_out, nsp[0] - mean_out) + wi * (nsp[1:] - mean_out).T @ (nsp[1:] - mean_out)
    cov_out = wc0 * np.outer(sp[0] - m, nsp[0] - mean_out) + wi * (sp[1:] - m).T @ (nsp[1:] - mean_out)
    return mean_out, var_out, cov_out


def mcf_moments(model, m, P, kw, Q, R, draws):
    """MCF.moment_approx (gaussfilt.py:186-198) with particles = m + draws, draws [NP,dx] ~ N(0,P) supplied."""
    func, cov = _step_fn(model, kw, Q, R)
    particles = m + draws
    tp = func(particles)
    n = draws.shape[0]
    mean_out = np.sum(tp, axis=0) / n
    var_out = cov + (tp - mean_out).T @ (tp - mean_out) / n
    cov_out = (particles - m).T @ (tp - mean_out) / n
    return mean_out, var_out, cov_out


def ekf_moments(model, m, P, kw, Q, R):
    """EKF.moment_approx (gaussfilt.py:217-252), first order (the second-order terms are commented out there)."""
    func, cov = _step_fn(model, kw, Q, R)
    J = (_Fx if kw == 'pred' else _Gx)(model, m)[0]
    return func(m)[0], cov + J @ P @ J.T, P @ J.T


def _gaussfilt_step(moments, m, P, y, draws_t=None):
    """one turn of GaussFilt.run / gaussfilt.GaussSumFilt.run: moment_approx('pred'), moment_approx('upd'), gain."""
    pm, pP, _ = moments(m, P, 'pred', None if draws_t is None else draws_t[0])
    mu_y, Sy, Cxy = moments(pm, pP, 'upd', None if draws_t is None else draws_t[1])
    K = Cxy @ np.linalg.inv(Sy)
    fm = pm + (y - mu_y) @ K.T
    fP = pP - K @ Sy @ K.T
    ll = gaussian_logpdf(np.reshape(y, [1, -1]), mu_y, Sy)
    return fm, fP, ll


def _moments_of(model, kind, Q, R, **kw):
    if kind == "EKF":
        return lambda m, P, step, d: ekf_moments(model, m, P, step, Q, R)
    if kind == "UKF":
        return lambda m, P, step, d: ukf_moments(model, m, P, step, Q, R, **kw)
    if kind == "MCF":
        return lambda m, P, step, d: mcf_moments(model, m, P, step, Q, R, d)
    raise ValueError(kind)


def gaussfilt_run(model, ys, Q, R, m0, P0, kind="EKF", draws=None, **kw):
    """GaussFilt.run (gaussfilt.py:88-130) over a moment approximation -> (ll [T], means [T,dx], covs [T,dx,dx]).
    draws [T,2,NP,dx] for the MCF (prediction draws ~ N(0, P_{t-1}), update draws ~ N(0, P_pred))."""
    mom = _moments_of(model, kind, Q, R, **kw)
    m, P = np.asarray(m0, np.float64).reshape(-1), np.asarray(P0, np.float64)
    ll, ms, Ps = [], [], []
    for t in range(ys.shape[0]):
        m, P, l = _gaussfilt_step(mom, m, P, ys[t], None if draws is None else draws[t])
        ll.append(l); ms.append(m); Ps.append(P)
    return np.asarray(ll), np.stack(ms), np.stack(Ps)


def bank_run(model, ys, Q, R, P0, init_means, kind="EKF", draws=None, **kw):
    """gaussfilt.GaussSumFilt.run (gaussfilt.py:325-361): weights multiply by exp(loglik) WITHOUT a max shift (:356)."""
    mom = _moments_of(model, kind, Q, R, **kw)
    T, (M, dx) = ys.shape[0], init_means.shape
    means = np.zeros((T + 1, dx, M)); covs = np.zeros((T + 1, dx, dx, M)); weights = np.zeros((T + 1, M))
    means[T] = init_means.T; covs[T] = np.asarray(P0, np.float64)[:, :, None]; weights[T] = 1.0 / M
    point = np.zeros((T, dx))
    for t in range(T):
        for m in range(M):
            fm, fP, l = _gaussfilt_step(mom, means[t - 1, :, m], covs[t - 1, :, :, m], ys[t],
                                        None if draws is None else draws[t, m])
            means[t, :, m] = fm; covs[t, :, :, m] = fP
            weights[t, m] = np.exp(l) * weights[t - 1, m]
        weights[t] /= np.sum(weights[t])
        point[t] = np.sum(np.multiply(means[t, :], weights[t]))
    return means, covs, weights, point
