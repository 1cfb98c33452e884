"""CPU oracle for the gaussfiltax filtering hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-NumPy restatement of the reference's JAX recursions
(`gaussfiltax/inference.py`, `gaussfiltax/utils.py`, `gaussfiltax/models.py`, the legacy NumPy classes).
It exists to *check* the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; nothing under
`bayesianfiltering_b200/` does, and the product path has no CPU fallback.

PARITY PINNED TO OUTPUTS OF THE REFERENCE ITSELF, RUN HERE.  The reference ships no test with an assertion
and no golden vector, and jax / tensorflow_probability are installed neither in the build container nor on the
GPU box.  `oracle/refshim.py` therefore registers NumPy-backed stand-ins under those module names,
`oracle/make_ref_fixtures.py` imports the UNMODIFIED `/root/reference/gaussfiltax` on top of them, runs its own
functions (`bootstrap_particle_filter`, `gaussian_sum_filter`, `unscented_gaussian_sum_filter`,
`speedy_augmented_gaussian_sum_filter`, `_kalman_step`, `optimal_resampling`, `NonlinearSSM.sample`, the legacy
classes) and freezes inputs + outputs as `tests/golden/ref_*.npz`; `tests/test_ref_pin.py` holds every oracle
function to those fixtures and `tests/test_gpu_ref.py` holds the CUDA kernels to them directly.

What the pin does NOT cover, because the third-party code is absent (refshim's header lists it): the arithmetic
of `jax.random.choice` (cumsum in `lax.associative_scan` order, `total*(1-u)`, fixed-bisection `searchsorted`),
TFP's `MultivariateNormalFullCovariance`, `jnp.linalg.solve` and forward-mode `jacfwd` is restated from the
published algorithms -- jax / tfp are not pinned by the reference either (`setup.py:7-10`; era evidence says
jax 0.4.1 .. 0.4.23) -- and threefry is not reproduced at all.  On top of the fixtures the oracle is held to
analytical known answers: the exact Kalman filter on linear-Gaussian models, Jacobians against float64 finite
differences, searchsorted against NumPy on sorted input, cumsum against float64.

Randomness is an *input* of every function here (initial particles, coloured process noise, resampling
uniforms, initial component means); the JAX PRNG is not reproduced.
"""
