"""GPU tier: the canonical prefix sum and the resampler are BIT-EXACT against the oracle (integer/index work)."""
import os

import numpy as np
import pytest

from oracle import scan as oscan

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _weights(n, seed, kind="pow"):
    rng = np.random.default_rng(seed)
    if kind == "pow":
        w = rng.random(n) ** 3
    elif kind == "sparse":
        w = rng.random(n) * (rng.random(n) < 0.05)
        w[rng.integers(n)] += 1.0
    elif kind == "onehot":
        w = np.zeros(n); w[rng.integers(n)] = 1.0
    elif kind == "uniform":
        w = np.ones(n)
    w = w.astype(np.float32)
    return (w / w.sum(dtype=np.float32)).astype(np.float32)


def _gpu_cumsum(w):
    import torch
    from bayesianfiltering_b200.engine import cumsum_canonical
    return cumsum_canonical(torch.from_numpy(w).cuda()).cpu().numpy()


def _gpu_anc(w, u, mode):
    import torch
    from bayesianfiltering_b200.engine import resample_ancestors
    return resample_ancestors(torch.from_numpy(w).cuda(), torch.from_numpy(np.atleast_1d(u)).cuda(), mode).cpu().numpy()


@pytest.mark.parametrize("n", [1, 2, 3, 5, 127, 128, 129, 1000, 4095, 4096, 4097, 8191, 8192, 10000, 65539, 1 << 20])
@pytest.mark.parametrize("kind", ["pow", "sparse"])
def test_cumsum_bit_exact(n, kind):
    w = _weights(n, n, kind)
    assert np.array_equal(_gpu_cumsum(w), oscan.cumsum_associative(w))


@pytest.mark.parametrize("n", [1, 2, 100, 4096, 4097, 10000, 65539, 1 << 20])
@pytest.mark.parametrize("kind", ["pow", "sparse", "onehot", "uniform"])
def test_ancestors_bit_exact_multinomial(n, kind):
    w = _weights(n, n + 1, kind)
    u = np.random.default_rng(n + 2).random(n, dtype=np.float32)
    if n > 4:
        u[0], u[1] = 0.0, np.float32(1.0 - 2 ** -24)        # extremes of [0,1)
    assert np.array_equal(_gpu_anc(w, u, "multinomial"), oscan.resample_ancestors(w, u, "multinomial"))


@pytest.mark.parametrize("n", [1, 7, 1000, 4097, 65539, 1 << 20])
@pytest.mark.parametrize("u0", [0.0, 0.37, 0.99999994])
def test_ancestors_bit_exact_systematic(n, u0):
    w = _weights(n, n + 3, "pow")
    a = _gpu_anc(w, np.float32(u0), "systematic")
    assert np.array_equal(a, oscan.resample_ancestors(w, np.float32(u0), "systematic"))
    assert np.all(np.diff(a) >= 0) or n > (1 << 16)       # monotone unless the fp32 cdf itself is not


def test_full_size_2pow24_bit_exact_and_checksum():
    """BASELINE config 3 size: N = 2^24.  Bit-exact cdf and ancestors against the oracle, plus the
    size-independent properties: counts sum to N, zero-weight particles never selected."""
    n = 1 << 24
    w = _weights(n, 24, "pow")
    w[::5] = 0.0
    cdf = _gpu_cumsum(w)
    assert np.array_equal(cdf, oscan.cumsum_associative(w))
    u = np.random.default_rng(25).random(n, dtype=np.float32)
    a = _gpu_anc(w, u, "multinomial")
    assert np.array_equal(a, oscan.resample_ancestors(w, u, "multinomial"))
    assert a.min() >= 0 and a.max() < n
    # at N = 2^24 the weights (~6e-8) are at the resolution of the fp32 cdf near 1, and the tree-ordered cdf is
    # not monotone in its last bit, so the reference's own search can land on a zero-weight slot; it stays rare
    assert np.mean(w[a] == 0.0) < 0.02
    s = _gpu_anc(w, np.float32(0.5), "systematic")
    assert np.array_equal(s, oscan.resample_ancestors(w, np.float32(0.5), "systematic"))
    assert np.bincount(s, minlength=n).sum() == n


def test_unnormalised_and_nan_weights_follow_the_oracle():
    rng = np.random.default_rng(5)
    w = (rng.random(5000) * 3.0).astype(np.float32)              # jr.choice scales r by cdf[-1]
    u = rng.random(5000, dtype=np.float32)
    assert np.array_equal(_gpu_anc(w, u, "multinomial"), oscan.resample_ancestors(w, u, "multinomial"))
    w[100] = np.nan
    g, o = _gpu_cumsum(w), oscan.cumsum_associative(w)
    assert np.array_equal(np.isnan(g), np.isnan(o)) and np.array_equal(g[~np.isnan(g)], o[~np.isnan(o)])


@pytest.mark.parametrize("name", ["scan_4097", "scan_10000"])
def test_golden_fixtures(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert np.array_equal(_gpu_cumsum(g["w"]), g["cdf"])
    assert np.array_equal(_gpu_anc(g["w"], g["u"], "multinomial"), g["anc_multinomial"])
    assert np.array_equal(_gpu_anc(g["w"], g["u"][:1], "systematic"), g["anc_systematic"])
