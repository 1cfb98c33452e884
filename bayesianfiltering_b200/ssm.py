"""State-space-model base class: the part of gaussfiltax/ssm.py:28-469 the hot path's callers use.

The reference's `SSM` is an abstract base copied from dynamax; of its methods only `sample` (and the
distribution accessors a subclass overrides) are ever called in the repository -- `fit_em` / `fit_sgd`
reference undefined helpers (ssm.py:381,459) and are dead code, so they are not reproduced.
"""
from abc import ABC


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

    def sample(self, params, key, num_timesteps, inputs=None):
        raise NotImplementedError
