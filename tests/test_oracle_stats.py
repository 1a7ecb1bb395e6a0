"""Closed-form regression oracle (oracle/daa_oracle.py) against independent numpy / scipy
evaluations of the same estimators.  statsmodels itself is absent ("parity unpinned")."""
import numpy as np
from scipy import stats as sps

from oracle import daa_oracle, philox


def _toy(n_val=2, N=12, C=3, J=9, R=5, seed=3):
    rng = np.random.default_rng(seed)
    scores = rng.standard_normal((n_val, N, J, C)).astype(np.float32)
    slope = rng.standard_normal((C, R)) * 0.3
    av = np.zeros((n_val, N, C, J, R), np.float32)
    for c in range(C):
        av[:, :, c] = (scores[..., c][..., None] * (slope[c] + 0.2 * rng.standard_normal((n_val, N, 1, R)))
                       + 0.1 * rng.standard_normal((n_val, N, J, R)))
    rec = rng.standard_normal((n_val, N, R)).astype(np.float32)
    return av, scores, rec


def test_hierarchical_matches_lstsq_and_ttest():
    av, sc, _ = _toy()
    p, coef, betas = daa_oracle.hierarchical_regression(av, sc)
    for v, c, r in [(0, 0, 0), (1, 2, 4), (0, 1, 3)]:
        b = []
        for g in range(av.shape[1]):
            x = sc[v, g, :, c].astype(np.float64)
            y = av[v, g, c, :, r].astype(np.float64)
            A = np.stack([np.ones_like(x), x], 1)       # "y ~ x": intercept + slope
            b.append(np.linalg.lstsq(A, y, rcond=None)[0][1])
        b = np.array(b)
        assert np.allclose(betas[v, c, :, r], b, rtol=1e-10, atol=1e-12)
        t, pv = sps.ttest_1samp(b, 0.0)                 # "beta ~ 1": intercept test
        assert np.isclose(coef[v, c, r], b.mean(), rtol=1e-12)
        assert np.isclose(p[v, c, r], pv, rtol=1e-9)


def test_fixed_matches_linregress():
    av, sc, rec = _toy()
    p, coef = daa_oracle.fixed_regression(av, sc, rec)
    for v, c, r in [(0, 0, 0), (1, 2, 4)]:
        x = sc[v, :, :, c].astype(np.float64).reshape(-1)
        y = (av[v, :, c, :, r].astype(np.float64) - rec[v, :, r].astype(np.float64)[:, None]).reshape(-1)
        lr = sps.linregress(x, y)
        assert np.isclose(coef[v, c, r], lr.slope, rtol=1e-10)
        assert np.isclose(p[v, c, r], lr.pvalue, rtol=1e-8)


def test_significance_vote():
    p = np.ones((4, 2, 3))
    thr = 0.05 / 3 / 2
    p[:3, 0, 1] = thr / 10          # 3 of 4 validations
    p[:2, 1, 2] = thr / 10          # 2 of 4
    s = daa_oracle.significant(p, 0.7)      # needs >= 2.8 votes
    assert s[0, 1] and not s[1, 2] and s.sum() == 1


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    f = lambda *a: [int(v) for v in philox.philox4x32_10(*a)]
    assert f(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f(*[0xffffffff] * 6) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normal_moments_and_slicing():
    z = philox.philox_normal(1037, philox.STREAM_DAA_AVATAR, 200000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    z2 = philox.philox_normal(1037, philox.STREAM_DAA_AVATAR, 1000, start=12345)
    assert np.array_equal(z2, z[12345:13345])       # sharding-invariant: pure function of index


def test_philox_rows_layout_and_slicing():
    """Row-addressed latent noise (DAA streams): sections start on Philox block boundaries, rows are
    independent of how many rows are drawn and where the slice starts."""
    L, styles = 20, [3, 20]
    ep = 20 + 4 + 20
    rows = philox.philox_rows(1037, philox.STREAM_DAA_AVATAR, 64, L, styles)
    assert rows.shape == (64, 43) and rows.dtype == np.float32
    flat = philox.philox_normal(1037, philox.STREAM_DAA_AVATAR, 64 * ep).reshape(64, ep)
    assert np.array_equal(rows[:, :20], flat[:, :20])            # content: blocks 0..4 of the row
    assert np.array_equal(rows[:, 20:23], flat[:, 20:23])        # style_0 (3 of its 4 padded draws)
    assert np.array_equal(rows[:, 23:], flat[:, 24:44])          # style_1 starts on the next block
    part = philox.philox_rows(1037, philox.STREAM_DAA_AVATAR, 10, L, styles, row_start=37)
    assert np.array_equal(part, rows[37:47])
    z = philox.philox_rows(5, philox.STREAM_DAA_BASE, 4096, L, styles)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
