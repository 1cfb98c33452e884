"""Minimal stand-in for the `jax.random` keys the reference's signatures take.

`jr.PRNGKey(seed)` in the reference is a uint32[2] array; ours is the same shape and feeds a Philox4x32-10
counter-based generator inside the kernels.  The JAX threefry stream is NOT reproduced: for bit-level
comparisons supply the random inputs explicitly (`noise=`, `resample_uniforms=`, `init_particles=`).
"""
import numpy as np


def PRNGKey(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def key_to_seed(key):
    if key is None:
        return 0
    if isinstance(key, (int, np.integer)):
        return int(key) & 0xFFFFFFFFFFFFFFFF
    k = np.asarray(key).astype(np.uint64).ravel()
    if k.size == 1:
        return int(k[0])
    return (int(k[0]) << 32 | int(k[1])) & 0xFFFFFFFFFFFFFFFF


def split(key, num=2):
    """Deterministic child keys (splitmix64 of the parent seed); same call shape as jr.split."""
    s = key_to_seed(key)
    out = []
    for i in range(num):
        z = (s + 0x9E3779B97F4A7C15 * (i + 1)) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
        out.append([z >> 32, z & 0xFFFFFFFF])
    return np.array(out, dtype=np.uint32)
