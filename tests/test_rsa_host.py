"""Host half of the RSA path (mopoe_b200/rsa.py: kendall_from_counts) against scipy.stats.kendalltau -- the function
the reference's fit_rsa calls (experiments/stat_utils.py:81-95) -- on the seven integers the CUDA kernel produces,
computed here by their definition (oracle/rsa_oracle.py: brute_counts)."""
import numpy as np
import pytest
from scipy.stats import kendalltau

from oracle import rsa_oracle as ro


def _check(x, y):
    from mopoe_b200 import rsa
    tau, p = rsa.kendall_from_counts(ro.brute_counts(x, y)[None], len(x))
    want = kendalltau(x, y)
    if np.isnan(want[0]):
        assert np.isnan(tau[0]) and np.isnan(p[0])
        return
    assert abs(tau[0] - want[0]) <= 1e-14
    assert abs(p[0] - want[1]) <= 1e-12 * max(want[1], 1e-300)


@pytest.mark.parametrize("size", [3, 5, 12, 33, 34, 120, 500])
def test_untied_vectors_exact_and_asymptotic_branches(size):
    rng = np.random.default_rng(size)
    x = rng.standard_normal(size)
    _check(x, x + rng.standard_normal(size))          # size <= 33: exact distribution; above: normal approximation
    _check(x, rng.standard_normal(size))
    _check(x, 2 * x + 1)                               # no discordant pair: exact branch at any size
    _check(x, -x)
    swapped = np.sort(x).copy()
    swapped[[0, 1]] = swapped[[1, 0]]
    _check(np.sort(x), swapped)                        # one discordant pair


@pytest.mark.parametrize("size", [6, 40, 400])
def test_tied_vectors_use_the_tie_corrected_variance(size):
    rng = np.random.default_rng(100 + size)
    x = np.round(rng.standard_normal(size) * 2)
    y = np.round(x + rng.standard_normal(size))
    _check(x, y)
    _check(x, rng.integers(0, 2, size).astype(float))  # a categorical reference
    _check(rng.standard_normal(size), y)               # ties on one side only
    _check(x, np.zeros(size))                          # constant reference: nan, nan


def test_rsa_table_shapes():
    rng = np.random.default_rng(1)
    lat = rng.standard_normal((15, 4)).astype(np.float32)
    sc = np.round(rng.standard_normal((15, 3))).astype(np.float32)
    cm, mats, kt = ro.rsa_table(lat, sc, {"age": rng.random(15), "sex": rng.integers(0, 2, 15)}, ["sex"])
    assert cm.shape == (15, 15) and mats.shape == (5, 15, 15) and kt.shape == (5, 2)
