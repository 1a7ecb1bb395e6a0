"""CPU restatement of the Digital Avatars Analysis sweep and its association statistics.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates experiments/workflow.py:361-427 (avatar generation), :452-513 (regression tables),
:517-537 (Bonferroni + vote) and experiments/stat_utils.py:55-79 (`make_regression`,
methods "hierarchical" and "fixed").  The model half runs through oracle/mopoe_oracle.forward,
which is pinned against the unmodified reference.  The statistics half is "parity unpinned":
the reference delegates to statsmodels (un-vendored, un-pinned, not installed); the closed forms
below are the textbook OLS / one-sample t-test that `sm.OLS.from_formula("y ~ x")` and
`sm.OLS.from_formula("beta ~ 1")` evaluate, cross-checked against numpy.linalg.lstsq and
scipy.stats in tests/test_oracle_stats.py.

Noise is injected (see oracle/philox.py for the production generator):
  eps_base  [n_val, M, N, E]                 the M stochastic reconstructions (workflow.py:388-398)
  eps_score [n_val, n_samples, N, n_scores]  Normal(loc_hat, scale_hat).sample (workflow.py:401-405)
  eps_av    [n_val, n_samples, n_scores, N, E]  one forward per (sample, score) (workflow.py:406-419)
"""
import numpy as np
import torch
from scipy import stats as sps

from . import mopoe_oracle as mo


def daa_generate(params, spec, src, dst, eps_base, eps_score, eps_av, sample_latents=True,
                 src_mod=0, dst_mod=1, given_scores=False):
    """src: (n_val, N, C) perturbed modality ("clinical"); dst: (n_val, N, R) read-out modality
    ("rois").  Returns float32 numpy arrays shaped like the reference's output files:
      avatars (n_val, N, C, J, R)   rois_digital_avatars.npy  (workflow.py:280-288,423-427)
      sampled_scores (n_val, N, J, C)  sampled_scores.npy      (workflow.py:420-427,435)
      reconstructions (n_val, N, R)    rois_reconstructions.npy (workflow.py:399-400,437)
    """
    sname, dname = spec.mod_names[src_mod], spec.mod_names[dst_mod]
    n_val, N, C = src.shape
    R = dst.shape[2]
    J = eps_score.shape[1]
    Mb = eps_base.shape[1]
    avatars = np.zeros((n_val, N, C, J, R), np.float32)
    sampled = np.zeros((n_val, N, J, C), np.float32)
    recons = np.zeros((n_val, N, R), np.float32)
    with torch.no_grad():
        for v in range(n_val):
            data = {sname: src[v], dname: dst[v]}
            loc_s, loc_d, scale_s = [], [], []
            for p in range(Mb):                                         # workflow.py:388-398
                rec = mo.forward(params, spec, data, eps_base[v, p], sample_latents=True)["rec"]
                loc_s.append(rec[sname][0])
                loc_d.append(rec[dname][0])
                scale_s.append(rec[sname][1].expand_as(rec[sname][0]))
            loc_hat = torch.stack(loc_s).mean(0)
            scale_hat = torch.stack(scale_s).mean(0)                    # (identical scales unless learn_output_sample_scale)
            recons[v] = torch.stack(loc_d).mean(0).numpy()
            scores = loc_hat + scale_hat * eps_score[v]                 # (J, N, C)  workflow.py:401-405
            if given_scores:                                            # sampling_strategy != "likelihood" (:337-346,
                scores = eps_score[v]                                   # :411-412): the values themselves
            for j in range(J):                                          # workflow.py:406-419
                for c in range(C):
                    cdata = src[v].clone()
                    cdata[:, c] = scores[j, :, c]
                    rec = mo.forward(params, spec, {sname: cdata, dname: dst[v]},
                                     eps_av[v, j, c] if sample_latents else None,
                                     sample_latents=sample_latents)["rec"]
                    avatars[v, :, c, j] = rec[dname][0].numpy()
            sampled[v] = scores.permute(1, 0, 2).numpy()                # swapaxes(0,1) workflow.py:420-422
    return avatars, sampled, recons


def hierarchical_regression(avatars, sampled_scores):
    """make_regression(method="hierarchical") (stat_utils.py:66-75) for every (val, score, roi).
    avatars (n_val,N,C,J,R) float32, sampled_scores (n_val,N,J,C) float32 -> float64
    pvalues, coefs (n_val,C,R) and per-subject slopes betas (n_val,C,N,R)."""
    n_val, N, C, J, R = avatars.shape
    pvalues = np.zeros((n_val, C, R))
    coefs = np.zeros((n_val, C, R))
    betas = np.zeros((n_val, C, N, R))
    for v in range(n_val):
        for c in range(C):
            x = sampled_scores[v, :, :, c].astype(np.float64)           # (N, J)
            y = avatars[v, :, c].astype(np.float64)                     # (N, J, R)
            xc = x - x.mean(1, keepdims=True)
            sxx = (xc * xc).sum(1)                                      # (N,)
            sxy = np.einsum("nj,njr->nr", xc, y - y.mean(1, keepdims=True))
            b = sxy / sxx[:, None]                                      # level 1: OLS slope per subject
            betas[v, c] = b
            mean = b.mean(0)                                            # level 2: beta ~ 1
            sd = b.std(0, ddof=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                t = mean / (sd / np.sqrt(N))
            coefs[v, c] = mean
            pvalues[v, c] = 2.0 * sps.t.sf(np.abs(t), N - 1)
    return pvalues, coefs, betas


def fixed_regression(avatars, sampled_scores, reconstructions):
    """make_regression(method="fixed") on roi_avatar_diff = avatar - reconstruction
    (workflow.py:489-492, stat_utils.py:62-63): one pooled simple OLS over N*J points."""
    n_val, N, C, J, R = avatars.shape
    pvalues = np.zeros((n_val, C, R))
    coefs = np.zeros((n_val, C, R))
    n = N * J
    for v in range(n_val):
        for c in range(C):
            x = sampled_scores[v, :, :, c].astype(np.float64).reshape(-1)
            y = (avatars[v, :, c].astype(np.float64)
                 - reconstructions[v].astype(np.float64)[:, None, :]).reshape(n, R)
            xc = x - x.mean()
            yc = y - y.mean(0)
            sxx = (xc * xc).sum()
            b = (xc[:, None] * yc).sum(0) / sxx
            rss = ((yc - xc[:, None] * b) ** 2).sum(0)
            se = np.sqrt(rss / (n - 2) / sxx)
            with np.errstate(divide="ignore", invalid="ignore"):
                t = b / se
            coefs[v, c] = b
            pvalues[v, c] = 2.0 * sps.t.sf(np.abs(t), n - 2)
    return pvalues, coefs


def significant(pvalues, trust_level):
    """workflow.py:517-523: Bonferroni threshold 0.05/n_rois/n_scores, vote over validations."""
    n_val, C, R = pvalues.shape
    thr = 0.05 / R / C
    return ((pvalues < thr).sum(axis=0) >= trust_level * n_val)


def significance_margin(pvalues):
    """min |log10 p - log10 thr| over the table: how far the closest decision is from flipping."""
    n_val, C, R = pvalues.shape
    thr = 0.05 / R / C
    with np.errstate(divide="ignore"):
        return float(np.min(np.abs(np.log10(np.maximum(pvalues, 1e-300)) - np.log10(thr))))
