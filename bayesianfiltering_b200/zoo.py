"""Compiled model registry, Python side.

The reference passes arbitrary Python callables `f(x, q, u)`, `h(x, r, u)` and a log-prob closure in its
parameter containers (gaussfiltax/models.py:46-50, :76-80) and traces them with JAX.  A CUDA kernel cannot
call back into Python, so the models the reference's drivers define are compiled into libbfilt.so
(csrc/models.cuh) and selected here: each constructor returns a `RegisteredModel` whose
`dynamics_function`, `emission_function` and `emission_log_prob` attributes are *tagged* callables to put
in `ParamsNLSSM` / `ParamsBPF` exactly where the reference's lambdas go.  The filter entry points read
the tag; an untagged callable raises (there is no CPU fallback).  Calling a tagged function evaluates it
on the GPU through `bf_model_eval`.

Sources of the definitions (reference): UNGM / L63 / SIN1D / LG -- docs/notebooks/Experiment_TSP_2023.ipynb
cell 2 (f3,g3 / lorentz_63,f63 / f1,g1 / fLG,gLG); BOT -- docs/experiments/BOT_Experiment_script.py:27-47.
"""
import numpy as np

from . import _lib


class _Tagged:
    """A registry function.  kind: 'f' (x,q,u)->x', 'h' (x,r,u)->y, 'lp' (x,y,u)->log p(y|x)."""

    def __init__(self, model, kind):
        self.__bf_model__ = model
        self.kind = kind
        self.__name__ = f"{model.name}_{kind}"

    def __call__(self, x, a=None, u=0.0, **kw):
        what = {"f": 0, "h": 1, "lp": 5}[self.kind]
        return self.__bf_model__.evaluate(what, x, a, u, **kw)

    def __repr__(self):
        return f"<bfilt registry function {self.__name__}>"


class RegisteredModel:
    def __init__(self, model_id, name, dx, dq, dy, dr, theta):
        self.model_id, self.name = model_id, name
        self.dx, self.dq, self.dy, self.dr = dx, dq, dy, dr
        self.theta = np.asarray(theta, dtype=np.float32).ravel()
        self.dynamics_function = _Tagged(self, "f")
        self.emission_function = _Tagged(self, "h")
        self.emission_log_prob = _Tagged(self, "lp")
        # single-argument faces for the legacy classes (gaussfilt.SSM(f=..., g=...)): f(x) = f(x, 0, 0), g(x) = h(x, 0, 0)
        self.f = _Tagged(self, "f")
        self.g = _Tagged(self, "h")

    # -- host-side description of a fully specified model (theta + the params' arrays)
    def host_model(self, q0=None, Q=None, r0=None, R=None, m0=None, P0=None):
        z = lambda v, n: np.zeros(n, np.float32) if v is None else v
        e = lambda v, n: np.eye(n, dtype=np.float32) if v is None else v
        return _lib.HostModel(self.model_id, self.dx, self.dq, self.dy, self.dr, self.theta,
                              z(q0, self.dq), e(Q, self.dq), z(r0, self.dr), e(R, self.dr), z(m0, self.dx), e(P0, self.dx))

    def evaluate(self, what, x, aux=None, u=0.0, R=None, r0=None, q0=None):
        """what: 0 f, 1 h, 2 dF/dx, 3 dF/dq, 4 dH/dx, 5 Gaussian emission log-prob of y=aux."""
        import torch
        lib = _lib.load()
        x_np = np.asarray(x.detach().cpu() if hasattr(x, "detach") else x, dtype=np.float32)
        single = x_np.ndim == 1
        x2 = np.ascontiguousarray(x_np.reshape(-1, self.dx))
        n = x2.shape[0]
        dev = torch.device("cuda", torch.cuda.current_device())
        xd = torch.from_numpy(x2).to(dev)
        auxd = None
        if aux is not None:
            a_np = np.asarray(aux.detach().cpu() if hasattr(aux, "detach") else aux, dtype=np.float32)
            if what in (0, 1):
                k = self.dq if what == 0 else self.dr
                a_np = np.full((1, k), a_np.item(), np.float32) if a_np.size == 1 else a_np.reshape(-1, k)
                a_np = np.ascontiguousarray(np.broadcast_to(a_np, (n, k)))
            else:
                a_np = np.ascontiguousarray(a_np.reshape(-1))
            auxd = torch.from_numpy(a_np).to(dev)
        shape = {0: (n, self.dx), 1: (n, self.dy), 2: (n, self.dx, self.dx), 3: (n, self.dx, self.dq),
                 4: (n, self.dy, self.dx), 5: (n,)}[what]
        out = torch.empty(shape, dtype=torch.float32, device=dev)
        hm = self.host_model(q0=q0, r0=r0, R=R)
        uu = float(np.asarray(u, dtype=np.float32).reshape(-1)[0]) if u is not None else 0.0
        _lib.check(lib.bf_model_eval(hm.desc, what, n, xd.data_ptr(), auxd.data_ptr() if auxd is not None else None, uu,
                                     out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        res = out.cpu().numpy()
        return res[0] if single else res

    def __repr__(self):
        return f"RegisteredModel({self.name}, dx={self.dx}, dq={self.dq}, dy={self.dy})"


def ungm(a=0.5, b=25.0, c=1.0, d=0.8, e=0.0):
    """f = a x + b x/(1+x^2) + c u + q ; h = d x + e x^2 + r  (the reference's "stochastic growth model")."""
    return RegisteredModel(_lib.MODEL_UNGM, "ungm", 1, 1, 1, 1, [a, b, c, d, e])


def bot(dt=0.5, acc=0.5, omega_scale=0.1, emission_dim=2):
    """Manoeuvring bearings-only tracking fManBOT with gBOT (emission_dim=1) or gBOT2 (emission_dim=2)."""
    if emission_dim not in (1, 2):
        raise ValueError("emission_dim must be 1 or 2")
    return RegisteredModel(_lib.MODEL_BOT, "bot", 4, 2, emission_dim, emission_dim, [dt, acc, omega_scale])


def lorenz63(sigma=10.0, rho=28.0, beta=2.667, dt=0.01, c=0.001):
    return RegisteredModel(_lib.MODEL_L63, "l63", 3, 3, 1, 1, [sigma, rho, beta, dt, c])


def lorenz96(dx=40, F=8.0, dt=0.01, c1=1.0, c2=0.05):
    """Lorenz-96, Euler step, every second state observed through c1 x + c2 x^2 (BASELINE configs[3]; the constants of
    gaussfiltax/nonlinearities.py:37-41 with the advection term restored).  Gaussian-sum filter only, on the
    large-state kernel (csrc/gsf_large.cuh); the noise covariances must be diagonal."""
    if dx % 2 or not 4 <= dx <= 64:
        raise ValueError("lorenz96 needs an even 4 <= dx <= 64")
    return RegisteredModel(_lib.MODEL_L96, "l96", dx, dx, dx // 2, dx // 2, [F, dt, c1, c2])


def linear_gaussian(A, H):
    A = np.asarray(A, dtype=np.float32); H = np.asarray(H, dtype=np.float32)
    dx, dy = A.shape[0], H.shape[0]
    if A.shape != (dx, dx) or H.shape != (dy, dx):
        raise ValueError("A must be [dx,dx] and H [dy,dx]")
    return RegisteredModel(_lib.MODEL_LG, "lg", dx, dx, dy, dy, np.concatenate([A.ravel(), H.ravel()]))


def sin1d(dx=1, w=10.0, c=1.0):
    return RegisteredModel(_lib.MODEL_SIN1D, "sin1d", dx, dx, 1, 1, [w, c])


def msv(phi, beta, sigma, h0):
    """Multivariate stochastic volatility with a switch (adaptive_experiment.py:47-57): f = Phi x + q,
    h = u beta exp(x/sigma) * r + (1-u)(H0 x + r) with diagonal Phi, H0 (vectors of length d <= 3).  The emission noise is
    multiplicative when u = 1; `.emission_log_prob` is the reference's `lmsvlp` with covariance M R M^T."""
    phi, beta, h0 = (np.atleast_1d(np.asarray(a, np.float32)) for a in (phi, beta, h0))
    d = phi.size
    if not (beta.size == d and h0.size == d and 1 <= d <= 3):
        raise ValueError("msv needs phi, beta, h0 of one length d <= 3")
    return RegisteredModel(_lib.MODEL_MSV, "msv", d, d, d, d, np.concatenate([phi, beta, [np.float32(sigma)], h0]))


def resolve(params):
    """The RegisteredModel behind a ParamsNLSSM / ParamsBPF, or raise: no untagged callables, no fallback."""
    f = getattr(params, "dynamics_function", None)
    h = getattr(params, "emission_function", None)
    mf = getattr(f, "__bf_model__", None)
    mh = getattr(h, "__bf_model__", None)
    if mf is None or mh is None:
        raise TypeError(
            "dynamics_function / emission_function must come from the compiled model registry "
            "(bayesianfiltering_b200.zoo); arbitrary Python callables cannot run inside the CUDA kernels "
            "and there is no CPU fallback")
    if mf is not mh:
        raise ValueError("dynamics_function and emission_function belong to different registry models")
    lp = getattr(params, "emission_distribution_log_prob", None)
    if lp is not None and getattr(lp, "__bf_model__", None) is not mf:
        raise TypeError("emission_distribution_log_prob must be the registry model's .emission_log_prob "
                        "(Gaussian with the params' emission_noise_bias / emission_noise_covariance)")
    return mf


def host_model_from_params(params, time_varying=False):
    """Registry model + its host description from a ParamsNLSSM / ParamsBPF.  With time_varying=True (Gaussian-sum
    filter) a noise covariance may carry a leading T axis (`_get_params`, inference.py:21): the description then holds
    its first matrix and the [T,d,d] arrays come back as the third / fourth value (None when constant)."""
    m = resolve(params)
    Q = np.asarray(params.dynamics_noise_covariance)
    R = np.asarray(params.emission_noise_covariance)
    for name in ("dynamics_noise_bias", "emission_noise_bias"):
        if np.asarray(getattr(params, name)).ndim > 1:        # the reference never time-indexes a bias (ndim 2 != dim + 1)
            raise NotImplementedError(f"{name} with a leading axis is not supported (the reference does not index it either)")
    Q_t = Q if Q.ndim == 3 else None
    R_t = R if R.ndim == 3 else None
    if (Q_t is not None or R_t is not None) and not time_varying:
        raise NotImplementedError("time-varying noise covariances (leading T axis, inference.py:21) are supported by "
                                  "gaussian_sum_filter only: the reference's particle filter fetches them but samples from "
                                  "params.* as given")
    hm = m.host_model(q0=params.dynamics_noise_bias, Q=(Q[0] if Q_t is not None else Q),
                      r0=params.emission_noise_bias, R=(R[0] if R_t is not None else R),
                      m0=params.initial_mean, P0=params.initial_covariance)
    return (m, hm, Q_t, R_t) if time_varying else (m, hm)
