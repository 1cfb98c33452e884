"""Augmented Gaussian-sum filter oracle (the "speedy" form) and optimal resampling.  TEST INFRASTRUCTURE ONLY.

Restates `gaussfiltax/inference.py:621-812` (`speedy_augmented_gaussian_sum_filter`) on top of `_predict` /
`_condition_on` (oracle/gsf.py) and `gaussfiltax/utils.py:216-244` (`optimal_resampling`).  Randomness is supplied:

  means0  [M,dx]          initial component means        (reference: MVN(m0,P0).sample(M, PRNGKey(0)), :795)
  z_iid   [M,dx,N1]       standard normals of branch 1   (reference: jr.normal(key, ...), :679 -- the SAME draw at
  s_iid   [M*N1,dx,N2]    standard normals of branch 2    every step, because `rng_key` is a closure constant, :677,:718)
  res_u   [M]             uniforms of jr.choice           (reference: jr.choice(PRNGKey(0), ...), :754 -- also the same every step)

A [T,...] leading axis on any of the three makes it time-varying (what a caller who threads keys would get).
"""
import numpy as np

from . import gsf as _gsf
from . import scan as _scan


def agsf_step(model, fm, fP, w, y, u, Q, q0, R, r0, z_iid, s_iid, res_u, opt_args):
    """One `_step` (inference.py:652-790).  fm [M,dx], fP [M,dx,dx], w [M] -> the same after the step."""
    dt = fm.dtype
    M, dx = fm.shape
    N1, N2 = z_iid.shape[2], s_iid.shape[2]
    a0, a1 = dt.type(opt_args[0]), dt.type(opt_args[1])
    Deltas = (a0 * fP).astype(dt)                                                   # :666
    L1 = np.linalg.cholesky(fP - Deltas).astype(dt)                                 # :686
    centered = (L1 @ z_iid.astype(dt)).astype(dt)                                   # :688  [M,dx,N1]
    z_sample = np.swapaxes(fm[:, :, None] + centered, 1, 2).astype(dt)              # :691-692  [M,N1,dx]
    # predict every (z_sample[m,j], Deltas[m])                                       :702-703
    zm = z_sample.reshape(M * N1, dx)
    zP = np.repeat(Deltas, N1, axis=0)
    pm, pP = _gsf.predict(model, zm, zP, Q, q0, u)                                  # [M*N1,dx], [M*N1,dx,dx]
    pw = (np.tile(w, (N1, 1)).T / dt.type(N1)).astype(dt).reshape(M * N1)           # :704, :709
    Lambdas = (a1 * pP).astype(dt)                                                  # :718
    L2 = np.linalg.cholesky(pP - Lambdas).astype(dt)                                # :728
    s_centered = (L2 @ s_iid.astype(dt)).astype(dt)                                 # :730  [M*N1,dx,N2]
    s_sample = np.swapaxes(pm[:, :, None] + s_centered, 1, 2).astype(dt)            # :733-734  [M*N1,N2,dx]
    sm = s_sample.reshape(M * N1 * N2, dx)
    sP = np.repeat(Lambdas, N2, axis=0)
    lls, um, uP = _gsf.condition_on(model, sm, sP, R, r0, u, y)                     # :741-742
    uw = (np.tile(pw, (N2, 1)).T / dt.type(N2)).astype(dt).reshape(M * N1 * N2)     # :743
    mx = np.max(lls)
    ls = np.exp(lls - mx).astype(dt)                                                # :745-746
    wu = (ls * uw).astype(dt)                                                       # :747
    pre_weights = (wu / np.sum(wu, dtype=dt)).astype(dt)                            # :748-749
    # jr.choice(PRNGKey(0), arange(L), shape=(M,), p=weights)                        :764
    idx = _scan.resample_ancestors(pre_weights, np.asarray(res_u, dtype=dt), "multinomial")
    nm, nP = um[idx], uP[idx]                                                       # :765-766
    nw = np.full(M, dt.type(1.0) / dt.type(M), dtype=dt)                            # :768
    return dict(means=nm, covariances=nP, weights=nw, pre_weights=pre_weights, updated_means=um, updated_covs=uP,
                ancestors=idx, lls=lls, loglik_inc=dt.type(mx + np.log(np.sum(wu, dtype=dt))))


def speedy_augmented_gaussian_sum_filter(model, emissions, inputs, means0, P0, Q, q0, R, r0, num_components, opt_args,
                                         z_iid, s_iid, res_u):
    """Whole filter; outputs [M,T,...] like the reference after `swap_axes_on_values` (:803)."""
    dt = means0.dtype
    T = emissions.shape[0]
    M, N1, N2 = (int(v) for v in num_components)
    fm = means0.copy()
    fP = np.broadcast_to(P0.astype(dt), (M,) + P0.shape).copy()                     # :796
    w = np.full(M, dt.type(1.0) / dt.type(M), dtype=dt)                             # :797
    tv = lambda a, nd, t: a[t] if np.ndim(a) == nd + 1 else a
    acc = dict(weights=[], means=[], covariances=[], pre_weights=[], updated_means=[], ancestors=[])
    for t in range(T):
        u = inputs[t] if inputs is not None else 0.0
        out = agsf_step(model, fm, fP, w, emissions[t], u, Q, q0, R, r0, tv(z_iid, 3, t), tv(s_iid, 3, t), tv(res_u, 1, t), opt_args)
        fm, fP, w = out["means"], out["covariances"], out["weights"]
        for k in acc:
            acc[k].append(out[k])
    res = {k: np.stack(v, axis=1) for k, v in acc.items() if k in ("weights", "means", "covariances")}
    res.update({k: np.stack(acc[k]) for k in ("pre_weights", "updated_means", "ancestors")})
    return res


def optimal_resampling(weights, N, u):
    """`utils.optimal_resampling` (utils.py:216-244; Fearnhead & Clifford 2003): keep the N - L largest weights, resample
    among the rest with the threshold weight p.  `u` [M]: the uniforms of its `jr.choice(key, M, (M,), p=res_weights)`.
    Returns (indices [N] int32, weights [N])."""
    w = np.asarray(weights)
    dt = w.dtype
    M = w.shape[0]
    sorted_idx = np.argsort(w, kind="stable")                                       # :224 (jnp.argsort is stable)
    sw = w[sorted_idx]                                                              # :223
    lower = np.triu(np.ones((M, M), dtype=dt), k=0).T                               # :227
    ps = (lower[M - N:M - 1] @ sw / np.arange(1, N).astype(dt)).astype(dt)          # :228
    ps = ps[::-1]                                                                   # :229
    ind = np.arange(1, N)
    b0, b1 = sw[M - ind - 1], sw[M - ind]                                           # :230
    preds = np.logical_and(b0 < ps, ps < b1)                                        # :231
    L = int(np.where(preds, ind, 0).sum())                                          # :232
    p = dt.type(1.0 / N) if L == 0 else ps[L - 1]                                   # :233
    rw = np.where(sw < p, sw, dt.type(0.0)).astype(dt)                              # :236
    rw = (rw / rw.sum(dtype=dt)).astype(dt)                                         # :237
    ridx = _scan.resample_ancestors(rw, np.asarray(u, dtype=dt), "multinomial")     # :238
    unsort = sorted_idx[ridx]                                                       # :239
    fidx = np.where(sw < p, unsort, sorted_idx)                                     # :241
    fw = np.where(sw < p, p, sw).astype(dt)                                         # :242
    return fidx[M - N:].astype(np.int32), (fw[M - N:] / fw[M - N:].sum(dtype=dt)).astype(dt)   # :244
