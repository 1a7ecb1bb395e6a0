"""CPU restatement of the RSA arithmetic (TEST INFRASTRUCTURE ONLY -- never imported by the product package).

Follows experiments/stat_utils.py of the reference line by line, calling the SAME third-party functions the
reference calls (scipy.spatial.distance.pdist / squareform, scipy.stats.kendalltau; environment.yml:15 pins
scipy=1.9.3, this image has a newer scipy with the same tau-b / "auto" p-value definition):
  data2cmat  stat_utils.py:25-33      cmat2triu  stat_utils.py:36-43
  vec2cmat   stat_utils.py:46-53      fit_rsa    stat_utils.py:81-95 (2-D branch)
and the per-(latent, score) loop of experiments/workflow.py:741-789 (rsa_exp).
Parity pinned: tests/test_oracle_vs_reference.py runs these against the unmodified reference module when
/root/reference is present."""
import numpy as np
from scipy.spatial.distance import pdist, squareform
from scipy.stats import kendalltau


def data2cmat(data):
    data = np.asarray(data)
    if data.ndim > 2:
        return np.array([squareform(pdist(d, metric="euclidean")) for d in data])
    return squareform(pdist(data, metric="euclidean"))


def cmat2triu(arr):
    assert np.ndim(arr) == 2 and arr.shape[0] == arr.shape[1]
    return arr[np.triu_indices(n=arr.shape[0], k=1)]


def vec2cmat(vec, categorical=False):
    vec = np.asarray(vec)
    if not categorical:
        return squareform(pdist(vec[:, None], metric="euclidean").transpose())
    return (vec[:, None] != vec).astype(int)


def fit_rsa(cmat, ref_cmat):
    tau, pval = kendalltau(cmat2triu(cmat), cmat2triu(ref_cmat))
    return tau, pval


def rsa_table(latents, scores, covariates, categorical):
    """workflow.py:757-779 for one latent block: latents (n, d), scores (n, C) float32, covariates {name: (n,)}.
    -> (cmat, scores_cmats (C + len(covariates), n, n), kendall (C + len(covariates), 2))."""
    cmat = data2cmat(latents)
    mats, out = [], []
    for c in range(scores.shape[1]):
        mats.append(vec2cmat(scores[:, c]))
        out.append(fit_rsa(cmat, mats[-1]))
    for name, v in covariates.items():
        mats.append(vec2cmat(v, categorical=name in categorical))
        out.append(fit_rsa(cmat, mats[-1]))
    return cmat, np.asarray(mats, dtype=np.float64), np.asarray(out, dtype=np.float64)


def brute_counts(x, y):
    """The seven integers of include/mopoe_b200.h: mopoe_rsa_kendall by definition (O(P^2) numpy, small P only)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    gx, gy = np.sign(x[:, None] - x[None, :]).astype(np.int64), np.sign(y[:, None] - y[None, :]).astype(np.int64)
    cx, cy = (gx == 0).sum(1) - 1, (gy == 0).sum(1) - 1
    return np.array([(gx * gy).sum(), cx.sum(), (cx * (cx - 1)).sum(), (cx * (2 * cx + 7)).sum(),
                     cy.sum(), (cy * (cy - 1)).sum(), (cy * (2 * cy + 7)).sum()], dtype=np.int64)
