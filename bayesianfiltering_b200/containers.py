"""Mixture containers, field-for-field as gaussfiltax/containers.py:17-61."""
from typing import Any, NamedTuple

import numpy as np


class GaussianComponent(NamedTuple):
    """Container object for a Gaussian mixture component (containers.py:17-23)."""
    mean: Any
    covariance: Any
    weight: Any


class GaussianSum(NamedTuple):
    """Container object for a Gaussian mixture (containers.py:32-41)."""
    means: Any
    covariances: Any
    weights: Any

    def _check_normalization(self):
        return bool(np.allclose(np.sum(np.asarray(self.weights)), 1.0))

    def _sum_weights(self):
        return np.sum(np.asarray(self.weights))


def _gaussian_sum_to_components(gaussian_sum: GaussianSum):
    """containers.py:43-44"""
    return [GaussianComponent(m, c, w) for m, c, w in zip(gaussian_sum.means, gaussian_sum.covariances, gaussian_sum.weights)]


def _components_to_gaussian_sum(components) -> GaussianSum:
    """containers.py:46-61"""
    return GaussianSum([c.mean for c in components], [c.covariance for c in components], [c.weight for c in components])
