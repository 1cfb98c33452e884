"""State-space-model base class: the part of gaussfiltax/ssm.py:28-469 the hot path's callers use.

The reference's `SSM` is an abstract base copied from dynamax.  Of its methods the repository only ever calls the
distribution accessors a subclass overrides and `sample` (ssm.py:169-211; `NonlinearGaussianSSM` inherits it,
`NonlinearSSM` overrides it with its own generator, which here runs on the GPU) -- `fit_em` / `fit_sgd` reference
undefined helpers (ssm.py:381,459) and are dead code, so they are not reproduced.
"""
from abc import ABC

import numpy as np


class SSM(ABC):
    @property
    def emission_shape(self):
        raise NotImplementedError

    @property
    def inputs_shape(self):
        return None

    def initial_distribution(self, params, inputs=None):
        raise NotImplementedError

    def transition_distribution(self, params, state, inputs=None):
        raise NotImplementedError

    def emission_distribution(self, params, state, inputs=None):
        raise NotImplementedError

    def log_prior(self, params):
        """ssm.py:140-150: no prior unless a subclass defines one."""
        return 0.0

    def sample(self, params, key, num_timesteps, inputs=None):
        """Sample states z_{1:T} and emissions y_{1:T} through the distribution accessors (ssm.py:169-211): the same
        order of draws and key splits as the reference's `lax.scan`, one step at a time on the host.  This is the generic
        path of a subclass that only defines the three distributions; `models.NonlinearSSM.sample` is the fused generator."""
        from .random import split
        key1, key2, key = split(key, 3)
        inp = (lambda t: None) if inputs is None else (lambda t: inputs[t])
        state = self.initial_distribution(params, inp(0)).sample(seed=key1)
        emission = self.emission_distribution(params, state, inp(0)).sample(seed=key2)
        states, emissions = [state], [emission]
        for t, k in enumerate(split(key, int(num_timesteps) - 1), start=1):
            k1, k2 = split(k, 2)
            state = self.transition_distribution(params, state, inp(t)).sample(seed=k2)
            emission = self.emission_distribution(params, state, inp(t)).sample(seed=k1)
            states.append(state); emissions.append(emission)
        return np.stack(states), np.stack(emissions)
