"""Bootstrap particle filter oracle.  TEST INFRASTRUCTURE ONLY.

Restates `gaussfiltax/inference.py:1302-1380` (`bootstrap_particle_filter`) with
`gaussfiltax/utils.py:207-214` (`_resample`) and `gaussfiltax/models.py:82-84`
(`ParamsBPF.sample_dynamics_distribution`).  Randomness is supplied:

  x0      [N,dx]     initial particles              (reference: MVN(m0,P0).sample, inference.py:1372-1373)
  noise   [T,N,dq]   coloured process noise q0+L_Q z (reference: MVN(q0,Q).sample, models.py:83)
  res_u   [T,N] or [T]  resampling uniforms          (reference: jr.choice's uniform draw, utils.py:210)
"""
import numpy as np

from . import scan as _scan


def pf_step(model, x, w, y, u, q, res_u, r0, R, ess_threshold=0.5, resampler="multinomial"):
    """One `_step` of the scan (inference.py:1330-1366).  Returns a dict with the post-resample carry
    and the by-products the CUDA path also reports (log-lik increment, ESS, ancestors)."""
    dt = x.dtype
    n = x.shape[0]
    # inference.py:1344-1345  new_particles = f(x_i, q_i, u)
    xt = model.f(x, q, u)
    # inference.py:1348-1349  lls = log_prob(new_particles, y, u)
    lls = model.emission_log_prob(xt, y, u, r0, R)
    # inference.py:1350-1353
    mx = np.max(lls)
    ls = np.exp(lls - mx).astype(dt)
    wu = (ls * w).astype(dt)
    z = np.sum(wu, dtype=dt)
    wn = (wu / z).astype(dt)
    # new output: log p(y_t | y_{1:t-1}) = max + log sum_i w_{t-1,i} exp(ll_i - max)
    loglik_inc = dt.type(mx + np.log(z))
    # inference.py:1356
    ess = dt.type(1.0) / np.sum(wn * wn, dtype=dt)
    do_resample = bool(ess < dt.type(ess_threshold) * dt.type(n))
    mean = np.sum(wn[:, None] * xt, axis=0, dtype=dt)      # point estimate of the drivers (BOT script :152)
    if do_resample:
        # utils.py:207-214
        anc = _scan.resample_ancestors(wn, res_u, resampler)
        x_out = xt[anc]
        w_out = np.full(n, dt.type(1.0) / dt.type(n), dtype=dt)
    else:
        anc = np.arange(n, dtype=np.int32)
        x_out, w_out = xt, wn
    return dict(particles=x_out, weights=w_out, propagated=xt, norm_weights=wn, lls=lls,
                loglik_inc=loglik_inc, ess=ess, resampled=do_resample, ancestors=anc, mean=mean)


def bootstrap_particle_filter(model, emissions, inputs, x0, noise, res_u, r0, R,
                              ess_threshold=0.5, resampler="multinomial", keep_trajectory=True):
    """The whole filter.  Outputs follow the reference's axis order after `swap_axes_on_values`
    (inference.py:1378): weights [N,T], particles [N,T,dx]; plus extras the reference does not return."""
    dt = x0.dtype
    T = emissions.shape[0]
    n = x0.shape[0]
    x = x0.copy()
    w = np.full(n, dt.type(1.0) / dt.type(n), dtype=dt)     # inference.py:1371
    W, X = [], []
    ll = np.zeros(T, dtype=dt); ess = np.zeros(T, dtype=dt); flag = np.zeros(T, dtype=np.int32)
    means = np.zeros((T, x0.shape[1]), dtype=dt)
    for t in range(T):
        u = inputs[t] if inputs is not None else 0.0
        ru = res_u[t] if res_u is not None else None
        out = pf_step(model, x, w, emissions[t], u, noise[t], ru, r0, R, ess_threshold, resampler)
        x, w = out["particles"], out["weights"]
        ll[t], ess[t], flag[t], means[t] = out["loglik_inc"], out["ess"], out["resampled"], out["mean"]
        if keep_trajectory:
            W.append(w); X.append(x)
    res = dict(loglik_inc=ll, marginal_loglik=np.sum(ll, dtype=np.float64), ess=ess, resampled=flag,
               means=means, final_particles=x, final_weights=w)
    if keep_trajectory:
        res["weights"] = np.stack(W, axis=1)
        res["particles"] = np.stack(X, axis=1)
    return res
