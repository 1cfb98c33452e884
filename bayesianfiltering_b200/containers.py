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


num_prt1 = 2      # containers.py:13-14: the branch counts the reference hard-codes for its tree version
num_prt2 = 2


def _branches_from_node1(node_component: GaussianComponent, splitting_cov, num_particles, key, *, normals=None):
    """containers.py:63-90: split a Gaussian component into `num_particles` Gaussian particles of covariance `splitting_cov`
    whose means are drawn from N(mean, covariance - splitting_cov); each gets weight / num_particles.  Host arithmetic (a
    handful of flops per component); inside the filters the same split runs in the kernel (csrc/agsf_kernels.cuh).
    `normals` [num_particles, dx] replaces the standard-normal draws behind jr.multivariate_normal."""
    from .random import key_to_seed
    mean = np.asarray(node_component.mean, np.float32)
    samp = np.asarray(node_component.covariance, np.float32) - np.asarray(splitting_cov, np.float32)
    n = int(num_particles)
    if normals is None:
        normals = np.random.default_rng(key_to_seed(key)).standard_normal((n, mean.shape[0]))
    with np.errstate(invalid="ignore"):
        try:
            L = np.linalg.cholesky(samp.astype(np.float64))
        except np.linalg.LinAlgError:
            L = np.full(samp.shape, np.nan)
    new_means = (mean + np.asarray(normals) @ L.T).astype(np.float32)
    new_means = np.where(np.isnan(new_means), mean, new_means)               # containers.py:85
    w = np.float32(node_component.weight) / np.float32(n)
    return [GaussianComponent(new_means[i], np.asarray(splitting_cov, np.float32), w) for i in range(n)]


_branches_from_node2 = _branches_from_node1          # containers.py:114-141: the same function with the second branch count


def _branches_from_tree1(components, split_covs_array, num_branch_array, key=0, *, normals=None):
    """containers.py:92-112: branch out every component of a mixture -> list (one per component) of lists of components.
    `normals` [num_components, num_branches, dx] replaces the standard-normal draws (one key per component otherwise)."""
    from .random import split
    keys = split(key, len(num_branch_array))
    zs = [None] * len(num_branch_array) if normals is None else normals
    return [_branches_from_node1(c, d, n, k, normals=z) for c, d, n, k, z in zip(components, split_covs_array, num_branch_array, keys, zs)]


_branches_from_tree2 = _branches_from_tree1          # containers.py:143-163
