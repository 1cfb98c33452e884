"""bayesianfiltering_b200 -- B200-native filtering recursions behind the `gaussfiltax` call surface.

Same module names as the reference package (`models`, `inference`, `containers`, `utils`, `ssm`,
`particlefilt`, `gausssumfilt`, `gaussfilt`), so `import gaussfiltax.inference as gf` becomes
`import bayesianfiltering_b200.inference as gf`.  The arithmetic lives in libbfilt.so (csrc/, hand-written
sm_100a CUDA behind the C ABI of include/bfilt.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (does not load the shared library until first use)

__version__ = "0.1.0"
