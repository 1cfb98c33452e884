"""CPU oracle for the gaussfiltax filtering hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-NumPy restatement of the reference's JAX recursions
(`gaussfiltax/inference.py`, `gaussfiltax/utils.py`, `gaussfiltax/models.py`).
It exists to *check* the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; nothing under
`bayesianfiltering_b200/` does, and the product path has no CPU fallback.

PARITY UNPINNED.  The reference cannot be imported in the build container (jax and
tensorflow_probability are not installed and there is no network), it ships no test
with an assertion, and its only stored numeric output (test_single_run.ipynb cell 9)
depends on the JAX threefry stream.  So this oracle is pinned only by
  * line-by-line correspondence with the reference sources (each function cites file:line),
  * the published algorithms of the third-party pieces it restates
    (jax.random.choice, lax.associative_scan / jnp.cumsum, jnp.searchsorted 'scan' method,
    TFP MultivariateNormalFullCovariance, jnp.linalg.solve) -- jax/tfp are unpinned by the
    reference (`setup.py:7-10`); era evidence says jax 0.4.1 .. 0.4.23,
  * analytical known answers we add: exact Kalman filter on linear-Gaussian models,
    Jacobians against float64 finite differences, searchsorted against numpy on sorted
    input, cumsum against float64.
Randomness is an *input* of every function here (initial particles, coloured process
noise, resampling uniforms, initial component means); the JAX PRNG is not reproduced.
"""
