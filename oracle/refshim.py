"""refshim -- TEST INFRASTRUCTURE ONLY.  A NumPy stand-in for the third-party runtime of the reference.

The reference (`/root/reference/gaussfiltax`) is pure Python on top of jax + tensorflow_probability (+ jaxtyping,
optax, fastprogress, matplotlib at import time); none of them is installed in the build container and there is no
network.  `install()` registers minimal modules under those names in `sys.modules`, all backed by
NumPy in float32, so that the UNMODIFIED reference source can be imported from where it lies and its own functions
(`inference.bootstrap_particle_filter`, `gaussian_sum_filter`, `unscented_gaussian_sum_filter`,
`speedy_augmented_gaussian_sum_filter`, `_kalman_step`, `utils._resample`, `optimal_resampling`,
`models.NonlinearSSM.sample`, the legacy NumPy classes) executed here.  `oracle/make_ref_fixtures.py` does that and
freezes inputs + outputs under `tests/golden/ref_*.npz`; `tests/test_ref_pin.py` holds the oracle to them.

What this pins and what it does not.  Executed as written: every line of the reference's own code on the path
(order of update / reweight / predict, `psd_solve`'s jitter, the ESS rule, `lax.cond` branch, axis swaps, sigma-point
formulas, the augmented filter's branching, the legacy classes' quirks).  RESTATED here, from the published
algorithms, because the third-party code itself is absent:
  * `jax.random.choice(key, a, shape, p=p)`  = `cumsum(p)` in `lax.associative_scan` order, `r = cdf[-1]*(1-u)`,
    `searchsorted(cdf, r)` by fixed bisection  (-> oracle/scan.py, the same restatement the kernels are tested on);
  * TFP `MultivariateNormalFullCovariance`   = `L = cholesky(cov)`; sample `loc + L z`; log_prob by forward
    substitution, `-0.5 z.z - sum(log diag L) - 0.5 k log(2 pi)`;
  * `jnp.linalg.solve` = LAPACK LU with partial pivoting (NumPy calls the same routine); `jacfwd` = central
    differences of the same Python callable evaluated in float64 (exact to ~1e-9, where jax's forward-mode AD is
    exact to fp32 rounding);
  * threefry is NOT reproduced: keys are turned into draws with NumPy's PCG64; every draw is recorded on `TAPE`
    so that the fixtures carry the random inputs the filter consumed (the kernels and the oracle take them as
    arguments).
"""
import importlib
import sys
import types

import numpy as np

from . import scan as _scan

REFERENCE_ROOT = "/root/reference"

# every random draw the reference code makes, in call order: (kind, payload dict)
TAPE = []
_X64 = [False]            # inside jacfwd: evaluate the user callables in float64


def _fdt():
    return np.float64 if _X64[0] else np.float32


def _cast(x):
    """jax's default precision: no float64 / int64 leaves a jnp call."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64 and not _X64[0]:
            return x.astype(np.float32)
        if x.dtype == np.int64:
            return x.astype(np.int32)
        return x
    if isinstance(x, np.floating):
        return x if _X64[0] else np.float32(x)
    if isinstance(x, np.integer) and not isinstance(x, np.bool_):
        return np.int32(x)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(_cast(v) for v in x)
    if isinstance(x, list):
        return [_cast(v) for v in x]
    return x


def _wrap(fn):
    def g(*a, **k):
        return _cast(fn(*a, **k))
    g.__name__ = getattr(fn, "__name__", "fn")
    return g


class _NS(types.ModuleType):
    """numpy namespace whose functions return jax-default dtypes."""

    def __init__(self, name, base, extra=None):
        super().__init__(name)
        self.__dict__["_base"] = base
        self.__dict__["_extra"] = extra or {}

    def __getattr__(self, name):
        ex = self.__dict__["_extra"]
        if name in ex:
            return ex[name]
        obj = getattr(self.__dict__["_base"], name)
        if callable(obj) and not isinstance(obj, type):
            return _wrap(obj)
        return obj


def _array(x, dtype=None):
    a = np.asarray(x, dtype=dtype)
    if dtype is None and a.dtype == np.float64:
        a = a.astype(_fdt())
    return _cast(a) if dtype is None else a


def _mk(fn):
    def g(shape, dtype=None):
        return fn(shape, dtype=_fdt() if dtype is None else dtype)
    return g


def _eye(n, m=None, k=0, dtype=None):
    return np.eye(n, m, k, dtype=_fdt() if dtype is None else dtype)


def _arange(*a, dtype=None):
    r = np.arange(*a, dtype=dtype)
    return _cast(r)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _S:
            def set(self, v):
                out = arr.copy(); out[idx] = v; return out

            def add(self, v):
                out = arr.copy(); out[idx] += v; return out
        return _S()


# ------------------------------------------------------------------------------------------ pytrees
def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(tree, is_leaf=None):
    leaves = []

    def rec(t):
        if is_leaf is not None and is_leaf(t):
            leaves.append(t); return ("leaf",)
        if t is None:
            return ("none",)
        if _is_namedtuple(t):
            return ("nt", type(t), [rec(v) for v in t])
        if isinstance(t, (list, tuple)):
            return ("seq", type(t), [rec(v) for v in t])
        if isinstance(t, dict):
            return ("dict", list(t.keys()), [rec(t[k]) for k in t])
        leaves.append(t)
        return ("leaf",)
    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        kind = d[0]
        if kind == "leaf":
            return next(it)
        if kind == "none":
            return None
        if kind == "nt":
            return d[1](*[rec(c) for c in d[2]])
        if kind == "seq":
            return d[1](rec(c) for c in d[2])
        return {k: rec(c) for k, c in zip(d[1], d[2])}
    return rec(treedef)


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = []
    for r in rest:
        # jax flattens the other trees only down to the structure of the first
        others.append(_flatten_up_to(td, r))
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


def _flatten_up_to(td, tree):
    out = []

    def rec(d, t):
        kind = d[0]
        if kind == "leaf":
            out.append(t)
        elif kind == "none":
            pass
        elif kind in ("nt", "seq"):
            for c, v in zip(d[2], list(t)):
                rec(c, v)
        else:
            for k, c in zip(d[1], d[2]):
                rec(c, t[k])
    rec(td, tree)
    return out


def tree_reduce(f, tree, initializer=None):
    import functools
    leaves, _ = tree_flatten(tree)
    return functools.reduce(f, leaves) if initializer is None else functools.reduce(f, leaves, initializer)


def _stack_trees(trees):
    leaves0, td = tree_flatten(trees[0])
    cols = [[] for _ in leaves0]
    for t in trees:
        lv, _ = tree_flatten(t)
        for c, v in zip(cols, lv):
            c.append(np.asarray(v))
    return tree_unflatten(td, [_cast(np.stack(c)) for c in cols])


# ------------------------------------------------------------------------------------------ transformations
def vmap(fun, in_axes=0, out_axes=0):
    """Loop over the mapped axis and stack: what jax.vmap computes, one element at a time."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                lv, _ = tree_flatten(a)
                n = np.shape(lv[0])[ax]
                break
        if n is None:
            raise ValueError("vmap needs at least one mapped argument")
        outs = []
        for i in range(n):
            call = []
            for a, ax in zip(args, axes):
                if ax is None:
                    call.append(a)
                else:
                    call.append(tree_map(lambda v: np.take(np.asarray(v), i, axis=ax), a))
            outs.append(fun(*call))
        res = _stack_trees(outs)
        if out_axes != 0:
            res = tree_map(lambda v: np.moveaxis(v, 0, out_axes), res)
        return res
    return mapped


def jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _jac(fun, argnums=0):
    """d fun / d arg[argnums] by float64 central differences of the same callable (jax: forward-mode AD)."""
    def J(*args):
        old = _X64[0]
        _X64[0] = True
        try:
            a64 = [np.asarray(a, dtype=np.float64) if isinstance(a, (np.ndarray, np.floating, float, list)) else a for a in args]
            x = np.array(a64[argnums], dtype=np.float64)
            y0 = np.asarray(fun(*a64), dtype=np.float64)
            jac = np.zeros(y0.shape + x.shape, dtype=np.float64)
            flat = x.reshape(-1)
            for k in range(flat.size):
                h = 1e-6 * max(1.0, abs(flat[k]))
                xp = flat.copy(); xp[k] += h
                xm = flat.copy(); xm[k] -= h
                a64[argnums] = xp.reshape(x.shape); yp = np.asarray(fun(*a64), dtype=np.float64)
                a64[argnums] = xm.reshape(x.shape); ym = np.asarray(fun(*a64), dtype=np.float64)
                jac.reshape(y0.shape + (flat.size,))[..., k] = (yp - ym) / (2 * h)
        finally:
            _X64[0] = old
        return jac if old else jac.astype(np.float32)
    return J


def scan(f, init, xs, length=None):
    carry = init
    ys = []
    lv, _ = tree_flatten(xs)
    n = length if xs is None else np.shape(lv[0])[0]
    for i in range(n):
        x = None if xs is None else tree_map(lambda v: np.asarray(v)[i], xs)
        carry, y = f(carry, x)
        ys.append(y)
    return carry, (_stack_trees(ys) if ys else None)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def while_loop(cond_fun, body_fun, init_val):
    v = init_val
    while bool(cond_fun(v)):
        v = body_fun(v)
    return v


def dynamic_slice(operand, start_indices, slice_sizes):
    idx = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start_indices, slice_sizes))
    return np.asarray(operand)[idx]


# ------------------------------------------------------------------------------------------ random
def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def _rng(key):
    k = np.asarray(key, dtype=np.uint32).reshape(-1)
    return np.random.Generator(np.random.PCG64([int(v) for v in k]))


def split(key, num=2):
    return _rng(key).integers(0, 2 ** 32, size=(int(num), 2), dtype=np.uint32)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    u = _rng(key).random(size=shape, dtype=np.float32)
    TAPE.append(("uniform", dict(u=u)))
    return _cast(minval + (maxval - minval) * u)


def normal(key, shape=(), dtype=None):
    z = _rng(key).standard_normal(size=shape).astype(np.float32)
    TAPE.append(("normal", dict(z=z)))
    return z


def multivariate_normal(key, mean, cov, shape=None, dtype=None, method="cholesky"):
    mean = np.asarray(mean, dtype=np.float32); cov = np.asarray(cov, dtype=np.float32)
    shp = tuple(shape) if shape is not None else ()
    z = _rng(key).standard_normal(size=shp + mean.shape).astype(np.float32)
    L = np.linalg.cholesky(cov).astype(np.float32)
    x = (mean + z @ L.T).astype(np.float32)
    TAPE.append(("multivariate_normal", dict(z=z, x=x)))
    return x


def choice(key, a, shape=(), replace=True, p=None, axis=0):
    """jax.random.choice with probabilities (jax/_src/random.py, replace=True branch):
        p_cuml = jnp.cumsum(p); r = p_cuml[-1] * (1 - uniform(key, shape)); ind = jnp.searchsorted(p_cuml, r)"""
    arr = np.arange(int(a)) if np.ndim(a) == 0 else np.asarray(a)
    shp = (shape,) if np.isscalar(shape) else tuple(shape)
    if p is None:
        idx = _rng(key).integers(0, arr.shape[0], size=shp)
        return _cast(arr[idx])
    p = np.asarray(p, dtype=np.float32)
    u = _rng(key).random(size=shp, dtype=np.float32)
    cdf = _scan.cumsum_associative(p)
    r = (cdf[-1] * (np.float32(1.0) - u)).astype(np.float32)
    idx = _scan.searchsorted_bisect(cdf, r.reshape(-1)).reshape(shp).astype(np.int32)
    TAPE.append(("choice", dict(p=p.copy(), u=u.copy(), idx=idx.copy())))
    return _cast(arr[idx])


# ------------------------------------------------------------------------------------------ TFP
class MultivariateNormalFullCovariance:
    """tfd.MultivariateNormalFullCovariance(loc, covariance_matrix): scale_tril = cholesky(cov)."""

    def __init__(self, loc=None, covariance_matrix=None, **kw):
        dt = _fdt()
        self.loc = np.asarray(loc, dtype=dt)
        self.cov = np.asarray(covariance_matrix, dtype=dt)
        self.L = np.linalg.cholesky(self.cov).astype(dt)

    def sample(self, sample_shape=(), seed=None):
        shp = (sample_shape,) if isinstance(sample_shape, (int, np.integer)) else tuple(sample_shape)
        z = _rng(seed).standard_normal(size=shp + self.loc.shape).astype(np.float32)
        x = (self.loc + z @ self.L.T).astype(np.float32)
        TAPE.append(("mvn_sample", dict(z=z, x=x)))
        return x

    def log_prob(self, y):
        dt = _fdt()
        d = (np.asarray(y, dtype=dt) - self.loc).astype(dt)
        k = self.loc.shape[-1]
        z = np.zeros_like(d)
        for i in range(k):                       # forward substitution, one row at a time
            s = d[..., i]
            for j in range(i):
                s = s - self.L[i, j] * z[..., j]
            z[..., i] = s / self.L[i, i]
        logdet = np.sum(np.log(np.diag(self.L)), dtype=dt)
        return (dt(-0.5) * np.sum(z * z, axis=-1, dtype=dt) - logdet - dt(0.5) * dt(k) * dt(np.log(2.0 * np.pi))).astype(dt) \
            if np.ndim(z) > 1 else dt(dt(-0.5) * np.sum(z * z, dtype=dt) - logdet - dt(0.5) * dt(k) * dt(np.log(2.0 * np.pi)))

    def mean(self):
        return self.loc

    def covariance(self):
        return self.cov


# ------------------------------------------------------------------------------------------ module assembly
class _Anything:
    """Stands in for names that are only imported / used in annotations (jaxtyping, matplotlib, optax, ...)."""

    def __init__(self, name="x"):
        self._n = name

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(self._n + "." + k)

    def __getitem__(self, k):
        return self

    def __call__(self, *a, **k):
        return self

    def __or__(self, o):
        return self

    def __mro_entries__(self, bases):
        return ()


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        return _Anything(self.__name__ + "." + k)


def _sqrtm(A):
    import scipy.linalg
    return scipy.linalg.sqrtm(np.asarray(A, dtype=np.float64))      # the reference takes jnp.real(...) of it


def install():
    """Register the stand-in modules (idempotent) and put the reference on sys.path."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "__refshim__", False):
        return
    for real in ("jax", "tensorflow_probability"):
        try:
            importlib.import_module(real)
            raise RuntimeError(f"a real `{real}` is importable here: use it instead of oracle/refshim.py")
        except ImportError:
            pass
    linalg = _NS("jax.numpy.linalg", np.linalg)
    jnp = _NS("jax.numpy", np, dict(
        array=_array, asarray=_array, zeros=_mk(np.zeros), ones=_mk(np.ones), empty=_mk(np.empty), eye=_eye,
        arange=_arange, linalg=linalg, float32=np.float32, int32=np.int32, float64=np.float64, ndarray=np.ndarray,
        pi=np.pi, inf=np.inf, newaxis=None, nan=np.nan, bool_=np.bool_, uint32=np.uint32,
        argsort=lambda a, axis=-1, **kw: _cast(np.argsort(a, axis=axis, kind="stable"))))      # jnp.argsort is stable
    jr = types.ModuleType("jax.random")
    for f in (PRNGKey, split, uniform, normal, multivariate_normal, choice):
        setattr(jr, f.__name__, f)
    jr.KeyArray = np.ndarray
    lax = types.ModuleType("jax.lax")
    lax.scan, lax.cond, lax.while_loop, lax.dynamic_slice = scan, cond, while_loop, dynamic_slice
    lax.stop_gradient = lambda x: x
    jtu = types.ModuleType("jax.tree_util")
    jtu.tree_flatten, jtu.tree_unflatten, jtu.tree_map, jtu.tree_reduce = tree_flatten, tree_unflatten, tree_map, tree_reduce
    jtu.register_pytree_node_class = lambda c: c
    jtu.tree_leaves = lambda t, is_leaf=None: tree_flatten(t, is_leaf)[0]
    jsl = types.ModuleType("jax.scipy.linalg")
    jsl.sqrtm = _sqrtm
    jsp = types.ModuleType("jax.scipy")
    jsp.linalg = jsl
    jsrc = types.ModuleType("jax._src")
    jsrc_random = types.ModuleType("jax._src.random")
    jsrc_random.KeyArray = np.ndarray
    jsrc.random = jsrc_random
    jax = types.ModuleType("jax")
    jax.__refshim__ = True
    jax.numpy, jax.random, jax.lax, jax.tree_util, jax.scipy, jax._src = jnp, jr, lax, jtu, jsp, jsrc
    jax.vmap, jax.pmap, jax.jit = vmap, vmap, jit
    jax.jacfwd = jax.jacrev = _jac
    jax.debug = _Anything("jax.debug")
    jax.device_put = lambda x, *a, **k: x
    jax.tree_map = tree_map
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": linalg, "jax.random": jr, "jax.lax": lax,
            "jax.tree_util": jtu, "jax.scipy": jsp, "jax.scipy.linalg": jsl, "jax._src": jsrc, "jax._src.random": jsrc_random}
    tfd = types.ModuleType("tensorflow_probability.substrates.jax.distributions")
    tfd.MultivariateNormalFullCovariance = MultivariateNormalFullCovariance
    tfd.Distribution = object
    tfb = _StubModule("tensorflow_probability.substrates.jax.bijectors")
    tfj = types.ModuleType("tensorflow_probability.substrates.jax")
    tfj.distributions, tfj.bijectors = tfd, tfb
    tfs = types.ModuleType("tensorflow_probability.substrates")
    tfs.jax = tfj
    tfp = types.ModuleType("tensorflow_probability")
    tfp.substrates = tfs
    mods.update({"tensorflow_probability": tfp, "tensorflow_probability.substrates": tfs,
                 "tensorflow_probability.substrates.jax": tfj,
                 "tensorflow_probability.substrates.jax.distributions": tfd,
                 "tensorflow_probability.substrates.jax.bijectors": tfb})
    for name in ("jaxtyping", "optax", "fastprogress", "fastprogress.fastprogress", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.animation", "matplotlib.patches"):
        if name not in sys.modules:
            mods[name] = _StubModule(name)
    sys.modules.update(mods)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def tape_clear():
    del TAPE[:]


def tape_take(kind):
    """Payloads of one kind, in call order."""
    return [p for k, p in TAPE if k == kind]
