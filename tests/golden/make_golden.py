"""Writes tests/golden/*.npz from the NumPy oracle.

The reference cannot be imported here (jax / tensorflow_probability are not installed) and it
ships no golden vectors, so these fixtures freeze OUR restatement of it: seeded inputs + every output of the
oracle for small cases.  The GPU parity tests compare the CUDA path against them as well as against the live
oracle, so the two cannot drift together unnoticed.   Regenerate:  python -m tests.golden.make_golden
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PF_CASES = {"pf_ungm": ("ungm", 40, 500, 0.5, "multinomial"), "pf_bot2": ("bot2", 30, 400, 1.0, "multinomial"),
            "pf_ungm_sys": ("ungm", 40, 500, 0.5, "systematic"), "pf_lg3": ("lg3", 25, 300, 0.5, "multinomial")}
GSF_CASES = {"gsf_ungm": ("ungm", 40, 12), "gsf_bot1": ("bot1", 30, 5), "gsf_lg42": ("lg42", 30, 3), "gsf_l63": ("l63", 30, 4)}
SCAN_CASES = {"scan_4097": 4097, "scan_10000": 10000}


def regenerate(name):
    from oracle import gsf as ogsf, pf as opf, scan as oscan
    from tests.parity_cases import case
    if name in PF_CASES:
        cname, T, N, thr, res = PF_CASES[name]
        c = case(cname)
        _, ys = c.simulate(T, 100)
        u = c.inputs(T)
        x0, noise, ru = c.pf_randomness(T, N, 101, res)
        o = opf.bootstrap_particle_filter(c.omodel, ys, u, x0, noise, ru, c.r0, c.R, thr, res)
        out = dict(emissions=ys, x0=x0, noise=noise, res_u=ru, weights=o["weights"], particles=o["particles"],
                   loglik_inc=o["loglik_inc"], ess=o["ess"], resampled=o["resampled"], means=o["means"])
        if u is not None:
            out["inputs"] = u
        return out
    if name in GSF_CASES:
        cname, T, M = GSF_CASES[name]
        c = case(cname)
        _, ys = c.simulate(T, 200)
        u = c.inputs(T)
        means0 = c.gsf_means0(M, 201)
        o = ogsf.gaussian_sum_filter(c.omodel, ys, u, means0, c.P0, c.Q, c.q0, c.R, c.r0)
        out = dict(emissions=ys, means0=means0, **{k: o[k] for k in ("weights", "means", "covariances", "predicted_means",
                                                                     "predicted_covariances", "loglik_inc")})
        if u is not None:
            out["inputs"] = u
        return out
    if name in SCAN_CASES:
        n = SCAN_CASES[name]
        rng = np.random.default_rng(n)
        w = (rng.random(n) ** 3).astype(np.float32)
        w /= w.sum(dtype=np.float32)
        u = rng.random(n, dtype=np.float32)
        return dict(w=w, u=u, cdf=oscan.cumsum_associative(w), anc_multinomial=oscan.resample_ancestors(w, u, "multinomial"),
                    anc_systematic=oscan.resample_ancestors(w, u[0], "systematic"))
    raise KeyError(name)


if __name__ == "__main__":
    for name in list(PF_CASES) + list(GSF_CASES) + list(SCAN_CASES):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **regenerate(name))
        print("wrote", name)
