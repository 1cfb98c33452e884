"""NumPy model zoo for the oracle.  TEST INFRASTRUCTURE ONLY.

Each model restates, vectorised over a leading particle/component axis, the `f(x, q, u)` /
`h(x, r, u)` lambdas the reference's drivers define (the reference takes arbitrary Python
callables; a CUDA kernel cannot, so the product keeps a compiled registry of these same
models -- `bayesianfiltering_b200/csrc/models.cuh`).  Sources:

  UNGM   "stochastic growth model"  docs/notebooks/Experiment_TSP_2023.ipynb cell 2 (f3, g3, g3lp)
  BOT    manoeuvring bearings-only  docs/experiments/BOT_Experiment_script.py:27-47 (fManBOT, gBOT, gBOT2)
  L63    Lorenz-63 Euler step       Experiment_TSP_2023.ipynb cell 2 (lorentz_63, f63; g1 = 0.001 x.x + r)
  L96    Lorenz-96                  gaussfiltax/nonlinearities.py:37-52 (constants; the advection term is
                                    restated as the true (x[i+1]-x[i-2])*x[i-1])
  LG     linear-Gaussian            docs/experiments/adaptive_experiment.py:59-64 (fLG, gLG)
  SIN1D  ICASSP toy                 Experiment_TSP_2023.ipynb cell 2 (f1 = sin(10x)+q, g1 = c x.x + r)

Jacobians are hand-derived (the reference uses `jacfwd`, inference.py:327-329); tests check them
against float64 central differences.  All arithmetic runs in the dtype of `x`.
"""
import numpy as np


def _f(dt, v):
    return np.asarray(v, dtype=dt)


class Model:
    name = "?"
    dx = dq = dy = dr = 0
    additive_q = True      # f(x,q,u) = f(x,0,u) + G q  with constant G
    additive_r = True      # h(x,r,u) = h(x,0,u) + r

    def theta(self):
        """Flat float32 parameter vector in the layout models.cuh expects."""
        raise NotImplementedError

    # --- vectorised over leading axis -------------------------------------------------
    def f(self, x, q, u):
        raise NotImplementedError

    def h(self, x, r, u):
        raise NotImplementedError

    def Fx(self, x, q, u):
        raise NotImplementedError

    def Fq(self, x, q, u):
        raise NotImplementedError

    def Hx(self, x, r, u):
        raise NotImplementedError

    def Hr(self, x, r, u):
        n = x.shape[0]
        return np.broadcast_to(np.eye(self.dy, self.dr, dtype=x.dtype), (n, self.dy, self.dr)).copy()

    def emission_log_prob(self, x, y, u, r0, R):
        """`MVN(loc=h(x, r0, u), covariance_matrix=R).log_prob(y)` -- the `*lp` lambdas of the drivers
        (e.g. BOT_Experiment_script.py:45).  TFP: L = chol(R); z = L^-1 (y - loc);
        -0.5 z.z - sum(log diag L) - 0.5 k log(2 pi)."""
        dt = x.dtype
        loc = self.h(x, np.broadcast_to(_f(dt, r0), (x.shape[0], self.dr)), u)
        return mvn_log_prob(loc, _f(dt, R), _f(dt, y))


def mvn_log_prob(loc, cov, y):
    """TFP MultivariateNormalFullCovariance(loc, cov).log_prob(y); loc [n,k], cov [k,k] or [n,k,k]."""
    dt = loc.dtype
    k = loc.shape[-1]
    L = np.linalg.cholesky(cov).astype(dt)
    d = (y - loc).astype(dt)
    if L.ndim == 2:
        z = np.linalg.solve(L, d.T).T.astype(dt)
        logdet = np.sum(np.log(np.diag(L))).astype(dt)
    else:
        z = np.linalg.solve(L, d[..., None])[..., 0].astype(dt)
        logdet = np.sum(np.log(np.diagonal(L, axis1=-2, axis2=-1)), axis=-1).astype(dt)
    half = dt.type(0.5)
    return (-half * np.sum(z * z, axis=-1) - logdet - half * dt.type(k) * dt.type(np.log(2.0 * np.pi))).astype(dt)


# ---------------------------------------------------------------------------------------
class UNGM(Model):
    """f = a x + b x/(1+x^2) + c u + q ; h = d x + e x^2 + r   (reference: a=.5 b=25 c=1 d=.8 e=0)."""
    name = "ungm"
    dx = dq = dy = dr = 1

    def __init__(self, a=0.5, b=25.0, c=1.0, d=0.8, e=0.0):
        self.a, self.b, self.c, self.d, self.e = a, b, c, d, e

    def theta(self):
        return np.array([self.a, self.b, self.c, self.d, self.e], dtype=np.float32)

    def f(self, x, q, u):
        dt = x.dtype
        a, b, c = dt.type(self.a), dt.type(self.b), dt.type(self.c)
        return (a * x + b * x / (dt.type(1) + x * x) + c * dt.type(u) + q).astype(dt)

    def h(self, x, r, u):
        dt = x.dtype
        return (dt.type(self.d) * x + dt.type(self.e) * x * x + r).astype(dt)

    def Fx(self, x, q, u):
        dt = x.dtype
        one = dt.type(1)
        x2 = x[:, 0] * x[:, 0]
        den = one + x2
        d = dt.type(self.a) + dt.type(self.b) * (one - x2) / (den * den)
        return d.astype(dt)[:, None, None]

    def Fq(self, x, q, u):
        return np.ones((x.shape[0], 1, 1), dtype=x.dtype)

    def Hx(self, x, r, u):
        dt = x.dtype
        return (dt.type(self.d) + dt.type(2.0 * self.e) * x[:, 0]).astype(dt)[:, None, None]


class BOT(Model):
    """Manoeuvring bearings-only tracking (BOT_Experiment_script.py:27-47).

    f = (c0 FCV + c1 FCT(x, acc) + c2 FCT(x, -acc)) x + G q,
        c0 = (u-1)(u-2)/2, c1 = -u(u-2), c2 = u(u-1)/2,  Omega(x,a) = w a / sqrt(x1^2 + x3^2)
    h = atan2(x2, x0) + r            (dy=1, gBOT)
    h = [atan2(x2, x0), hypot] + r   (dy=2, gBOT2)
    """
    name = "bot"
    dx, dq = 4, 2

    def __init__(self, dt=0.5, acc=0.5, w=0.1, dy=2):
        self.dt, self.acc, self.w = dt, acc, w
        self.dy = self.dr = dy

    def theta(self):
        return np.array([self.dt, self.acc, self.w], dtype=np.float32)

    def _coef(self, dt_, u):
        u = dt_.type(u)
        one, two, half = dt_.type(1), dt_.type(2), dt_.type(0.5)
        return half * (u - one) * (u - two), -u * (u - two), half * u * (u - one)

    def _fct_apply(self, x, a):
        """FCT(x, a) @ x and the pieces needed for the Jacobian."""
        dt_ = x.dtype
        T = dt_.type(self.dt)
        s = np.sqrt(x[:, 1] * x[:, 1] + x[:, 3] * x[:, 3])
        om = dt_.type(self.w) * dt_.type(a) / s
        sn, cs = np.sin(T * om), np.cos(T * om)
        A = sn / om
        Bc = (dt_.type(1) - cs) / om
        out = np.stack([x[:, 0] + A * x[:, 1] - Bc * x[:, 3],
                        cs * x[:, 1] - sn * x[:, 3],
                        Bc * x[:, 1] + x[:, 2] + A * x[:, 3],
                        sn * x[:, 1] + cs * x[:, 3]], axis=1)
        return out.astype(dt_), (s, om, sn, cs, A, Bc)

    def f(self, x, q, u):
        dt_ = x.dtype
        T = dt_.type(self.dt)
        c0, c1, c2 = self._coef(dt_, u)
        cv = np.stack([x[:, 0] + T * x[:, 1], x[:, 1], x[:, 2] + T * x[:, 3], x[:, 3]], axis=1)
        ct1, _ = self._fct_apply(x, self.acc)
        ct2, _ = self._fct_apply(x, -self.acc)
        half = dt_.type(0.5)
        gq = np.stack([half * q[:, 0], q[:, 0], half * q[:, 1], q[:, 1]], axis=1)
        return (c0 * cv + c1 * ct1 + c2 * ct2 + gq).astype(dt_)

    def _fct_jac(self, x, a):
        dt_ = x.dtype
        T = dt_.type(self.dt)
        n = x.shape[0]
        _, (s, om, sn, cs, A, Bc) = self._fct_apply(x, a)
        one = dt_.type(1)
        # d/dOmega of the four entries
        dA = (T * cs * om - sn) / (om * om)
        dB = (T * sn * om - (one - cs)) / (om * om)
        dsn = T * cs
        dcs = -T * sn
        # d(FCT x)/dOmega
        v = np.stack([dA * x[:, 1] - dB * x[:, 3],
                      dcs * x[:, 1] - dsn * x[:, 3],
                      dB * x[:, 1] + dA * x[:, 3],
                      dsn * x[:, 1] + dcs * x[:, 3]], axis=1)
        # grad Omega: dOm/dx1 = -Om x1 / s^2, dOm/dx3 = -Om x3 / s^2
        g1 = -om * x[:, 1] / (s * s)
        g3 = -om * x[:, 3] / (s * s)
        J = np.zeros((n, 4, 4), dtype=dt_)
        J[:, 0, 0] = one; J[:, 0, 1] = A; J[:, 0, 3] = -Bc
        J[:, 1, 1] = cs; J[:, 1, 3] = -sn
        J[:, 2, 1] = Bc; J[:, 2, 2] = one; J[:, 2, 3] = A
        J[:, 3, 1] = sn; J[:, 3, 3] = cs
        J[:, :, 1] += v * g1[:, None]
        J[:, :, 3] += v * g3[:, None]
        return J

    def Fx(self, x, q, u):
        dt_ = x.dtype
        T = dt_.type(self.dt)
        c0, c1, c2 = self._coef(dt_, u)
        FCV = np.array([[1, T, 0, 0], [0, 1, 0, 0], [0, 0, 1, T], [0, 0, 0, 1]], dtype=dt_)
        return (c0 * FCV[None] + c1 * self._fct_jac(x, self.acc) + c2 * self._fct_jac(x, -self.acc)).astype(dt_)

    def Fq(self, x, q, u):
        G = np.array([[0.5, 0], [1, 0], [0, 0.5], [0, 1]], dtype=x.dtype)
        return np.broadcast_to(G, (x.shape[0], 4, 2)).copy()

    def h(self, x, r, u):
        dt_ = x.dtype
        b = np.arctan2(x[:, 2], x[:, 0])
        if self.dy == 1:
            return (b[:, None] + r).astype(dt_)
        rng = np.sqrt(x[:, 0] * x[:, 0] + x[:, 2] * x[:, 2])
        return (np.stack([b, rng], axis=1) + r).astype(dt_)

    def Hx(self, x, r, u):
        dt_ = x.dtype
        n = x.shape[0]
        d2 = x[:, 0] * x[:, 0] + x[:, 2] * x[:, 2]
        H = np.zeros((n, self.dy, 4), dtype=dt_)
        H[:, 0, 0] = -x[:, 2] / d2
        H[:, 0, 2] = x[:, 0] / d2
        if self.dy == 2:
            rng = np.sqrt(d2)
            H[:, 1, 0] = x[:, 0] / rng
            H[:, 1, 2] = x[:, 2] / rng
        return H


class L63(Model):
    """Euler-discretised Lorenz-63, h = c x.x + r (Experiment_TSP_2023.ipynb cell 2)."""
    name = "l63"
    dx = dq = 3
    dy = dr = 1

    def __init__(self, sigma=10.0, rho=28.0, beta=2.667, dt=0.01, c=0.001):
        self.sigma, self.rho, self.beta, self.dt, self.c = sigma, rho, beta, dt, c

    def theta(self):
        return np.array([self.sigma, self.rho, self.beta, self.dt, self.c], dtype=np.float32)

    def f(self, x, q, u):
        t = x.dtype.type
        dt_, sg, rho, be = t(self.dt), t(self.sigma), t(self.rho), t(self.beta)
        dxx = dt_ * sg * (x[:, 1] - x[:, 0])
        dyy = dt_ * (x[:, 0] * rho - x[:, 1] - x[:, 0] * x[:, 2])
        dzz = dt_ * (x[:, 0] * x[:, 1] - be * x[:, 2])
        return (np.stack([dxx + x[:, 0], dyy + x[:, 1], dzz + x[:, 2]], axis=1) + q).astype(x.dtype)

    def Fx(self, x, q, u):
        t = x.dtype.type
        dt_, sg, rho, be = t(self.dt), t(self.sigma), t(self.rho), t(self.beta)
        n = x.shape[0]
        J = np.zeros((n, 3, 3), dtype=x.dtype)
        J[:, 0, 0] = t(1) - dt_ * sg; J[:, 0, 1] = dt_ * sg
        J[:, 1, 0] = dt_ * (rho - x[:, 2]); J[:, 1, 1] = t(1) - dt_; J[:, 1, 2] = -dt_ * x[:, 0]
        J[:, 2, 0] = dt_ * x[:, 1]; J[:, 2, 1] = dt_ * x[:, 0]; J[:, 2, 2] = t(1) - dt_ * be
        return J

    def Fq(self, x, q, u):
        return np.broadcast_to(np.eye(3, dtype=x.dtype), (x.shape[0], 3, 3)).copy()

    def h(self, x, r, u):
        return (x.dtype.type(self.c) * np.sum(x * x, axis=1, keepdims=True) + r).astype(x.dtype)

    def Hx(self, x, r, u):
        return (x.dtype.type(2.0 * self.c) * x)[:, None, :].astype(x.dtype)


class L96(Model):
    """Lorenz-96, Euler step: x_i += dt((x_{i+1} - x_{i-2}) x_{i-1} - x_i + F) + q_i.
    h_j = c1 x_{2j} + c2 x_{2j}^2 + r_j, j < dx/2 (nonlinearities.py:37-52 observes every 2nd state)."""
    name = "l96"

    def __init__(self, dx=40, F=8.0, dt=0.01, c1=1.0, c2=0.05):
        self.dx = self.dq = dx
        self.dy = self.dr = dx // 2
        self.F, self.dt, self.c1, self.c2 = F, dt, c1, c2

    def theta(self):
        return np.array([self.F, self.dt, self.c1, self.c2], dtype=np.float32)

    def f(self, x, q, u):
        t = x.dtype.type
        xp1 = np.roll(x, -1, axis=1); xm1 = np.roll(x, 1, axis=1); xm2 = np.roll(x, 2, axis=1)
        return (x + t(self.dt) * ((xp1 - xm2) * xm1 - x + t(self.F)) + q).astype(x.dtype)

    def Fx(self, x, q, u):
        t = x.dtype.type
        n, d = x.shape
        dt_ = t(self.dt)
        J = np.zeros((n, d, d), dtype=x.dtype)
        idx = np.arange(d)
        xp1 = np.roll(x, -1, axis=1); xm1 = np.roll(x, 1, axis=1); xm2 = np.roll(x, 2, axis=1)
        J[:, idx, idx] = t(1) - dt_
        J[:, idx, (idx + 1) % d] += dt_ * xm1
        J[:, idx, (idx - 2) % d] += -dt_ * xm1
        J[:, idx, (idx - 1) % d] += dt_ * (xp1 - xm2)
        return J

    def Fq(self, x, q, u):
        return np.broadcast_to(np.eye(self.dx, dtype=x.dtype), (x.shape[0], self.dx, self.dx)).copy()

    def h(self, x, r, u):
        t = x.dtype.type
        xs = x[:, 0:2 * self.dy:2]
        return (t(self.c1) * xs + t(self.c2) * xs * xs + r).astype(x.dtype)

    def Hx(self, x, r, u):
        t = x.dtype.type
        n = x.shape[0]
        H = np.zeros((n, self.dy, self.dx), dtype=x.dtype)
        j = np.arange(self.dy)
        H[:, j, 2 * j] = t(self.c1) + t(2.0 * self.c2) * x[:, 0:2 * self.dy:2]
        return H


class LG(Model):
    """f = A x + q ; h = H x + r (adaptive_experiment.py:59-64)."""
    name = "lg"

    def __init__(self, A, H):
        self.A = np.asarray(A, dtype=np.float64)
        self.H = np.asarray(H, dtype=np.float64)
        self.dx = self.dq = self.A.shape[0]
        self.dy = self.dr = self.H.shape[0]

    def theta(self):
        return np.concatenate([self.A.ravel(), self.H.ravel()]).astype(np.float32)

    def f(self, x, q, u):
        return (x @ self.A.astype(x.dtype).T + q).astype(x.dtype)

    def h(self, x, r, u):
        return (x @ self.H.astype(x.dtype).T + r).astype(x.dtype)

    def Fx(self, x, q, u):
        return np.broadcast_to(self.A.astype(x.dtype), (x.shape[0], self.dx, self.dx)).copy()

    def Fq(self, x, q, u):
        return np.broadcast_to(np.eye(self.dx, dtype=x.dtype), (x.shape[0], self.dx, self.dx)).copy()

    def Hx(self, x, r, u):
        return np.broadcast_to(self.H.astype(x.dtype), (x.shape[0], self.dy, self.dx)).copy()


class SIN1D(Model):
    """f = sin(w x) + q (element-wise) ; h = c x.x + r (Experiment A / ICASSP toy)."""
    name = "sin1d"
    dy = dr = 1

    def __init__(self, dx=1, w=10.0, c=1.0):
        self.dx = self.dq = dx
        self.w, self.c = w, c

    def theta(self):
        return np.array([self.w, self.c], dtype=np.float32)

    def f(self, x, q, u):
        return (np.sin(x.dtype.type(self.w) * x) + q).astype(x.dtype)

    def Fx(self, x, q, u):
        t = x.dtype.type
        n = x.shape[0]
        J = np.zeros((n, self.dx, self.dx), dtype=x.dtype)
        i = np.arange(self.dx)
        J[:, i, i] = t(self.w) * np.cos(t(self.w) * x)
        return J

    def Fq(self, x, q, u):
        return np.broadcast_to(np.eye(self.dx, dtype=x.dtype), (x.shape[0], self.dx, self.dx)).copy()

    def h(self, x, r, u):
        return (x.dtype.type(self.c) * np.sum(x * x, axis=1, keepdims=True) + r).astype(x.dtype)

    def Hx(self, x, r, u):
        return (x.dtype.type(2.0 * self.c) * x)[:, None, :].astype(x.dtype)


class MSV(Model):
    """Multivariate stochastic volatility with a switch (adaptive_experiment.py:47-57, `fmsv`, `glmsv`, `lmsvlp`):
        f = Phi x + q                                   (Phi diagonal)
        h = u beta exp(x/sigma) * r + (1-u) (H0 x + r)  (H0 diagonal; the emission noise enters MULTIPLICATIVELY when u = 1)
        log p(y|x) = MVN(h(x, r0, u), M R M^T).log_prob(y),  M = u diag(beta exp(x/sigma)) + (1-u) I
    theta = [Phi diag (d) | beta (d) | sigma | H0 diag (d)], dx = dq = dy = dr = d."""
    name = "msv"
    additive_r = False

    def __init__(self, phi, beta, sigma, h0):
        self.phi, self.beta, self.h0 = (np.atleast_1d(np.asarray(a, np.float64)) for a in (phi, beta, h0))
        self.sigma = float(sigma)
        self.dx = self.dq = self.dy = self.dr = self.phi.shape[0]

    def theta(self):
        return np.concatenate([self.phi, self.beta, [self.sigma], self.h0]).astype(np.float32)

    def f(self, x, q, u):
        return (self.phi.astype(x.dtype) * x + q).astype(x.dtype)

    def Fx(self, x, q, u):
        return np.broadcast_to(np.diag(self.phi).astype(x.dtype), (x.shape[0], self.dx, self.dx)).copy()

    def Fq(self, x, q, u):
        return np.broadcast_to(np.eye(self.dx, dtype=x.dtype), (x.shape[0], self.dx, self.dx)).copy()

    def _d(self, x, u):
        """diagonal of M = dh/dr"""
        t = x.dtype.type
        return (t(u) * self.beta.astype(x.dtype) * np.exp(x / t(self.sigma)) + (t(1) - t(u))).astype(x.dtype)

    def h(self, x, r, u):
        t = x.dtype.type
        return (t(u) * self.beta.astype(x.dtype) * np.exp(x / t(self.sigma)) * r
                + (t(1) - t(u)) * (self.h0.astype(x.dtype) * x + r)).astype(x.dtype)

    def Hx(self, x, r, u):
        t = x.dtype.type
        d = t(u) * self.beta.astype(x.dtype) * np.exp(x / t(self.sigma)) * r / t(self.sigma) + (t(1) - t(u)) * self.h0.astype(x.dtype)
        J = np.zeros((x.shape[0], self.dy, self.dx), dtype=x.dtype)
        i = np.arange(self.dx)
        J[:, i, i] = d
        return J

    def Hr(self, x, r, u):
        J = np.zeros((x.shape[0], self.dy, self.dr), dtype=x.dtype)
        i = np.arange(self.dx)
        J[:, i, i] = self._d(x, u)
        return J

    def emission_log_prob(self, x, y, u, r0, R):
        dt = x.dtype
        n = x.shape[0]
        loc = self.h(x, np.broadcast_to(_f(dt, r0), (n, self.dr)), u)
        d = self._d(x, u)
        cov = (d[:, :, None] * _f(dt, R)[None] * d[:, None, :]).astype(dt)       # M R M^T, state-dependent
        return mvn_log_prob(loc, cov, _f(dt, y))


def numerical_jacobian(fun, x, eps=1e-6):
    """float64 central differences of fun: [n,d] -> [n,k]; returns [n,k,d]."""
    x = np.asarray(x, dtype=np.float64)
    n, d = x.shape
    cols = []
    for j in range(d):
        e = np.zeros(d); e[j] = eps
        cols.append((fun(x + e) - fun(x - e)) / (2 * eps))
    return np.stack(cols, axis=-1)
