"""Mirror of experiments/stat_utils.make_regression (stat_utils.py:55-79) on the B200 path.

The reference fits one statsmodels model per (validation, score, ROI) from a pandas frame; here the
same estimators are evaluated for a whole frame of series on the GPU (csrc/mopoe_daa.cu).
`make_regression` keeps the reference signature for a single series."""
import numpy as np
import pandas as pd
import torch

from . import daa


def make_regression(df, x_name, y_name, other_cov_names=[], groups_name=None, method="fixed", other=None):
    """-> (pvalue, coef, per-subject betas DataFrame | None), like stat_utils.py:55-79.
    Supported: method "hierarchical" (per-group OLS slope, then one-sample t-test of the slopes) and
    "fixed" (pooled simple OLS); no extra covariates.  "mixed" (MixedLM) is not on this path."""
    if other_cov_names:
        raise NotImplementedError("other_cov_names is not on the B200 path")
    if method not in ("hierarchical", "fixed"):
        raise NotImplementedError("method=%r is not on the B200 path (hierarchical, fixed)" % (method,))
    if method == "hierarchical":
        groups = list(dict.fromkeys(df[groups_name].tolist()))       # groupby(sort=False) order
        sizes = {len(df[df[groups_name] == g]) for g in groups}
        if len(sizes) != 1:
            raise NotImplementedError("groups of unequal size are not on the B200 path")
        J = sizes.pop()
        x = np.stack([df.loc[df[groups_name] == g, x_name].to_numpy(np.float32) for g in groups])
        y = np.stack([df.loc[df[groups_name] == g, y_name].to_numpy(np.float32) for g in groups])
    else:
        groups = [0]
        x = df[x_name].to_numpy(np.float32)[None]
        y = df[y_name].to_numpy(np.float32)[None]
        J = x.shape[1]
    N = len(groups)
    if method == "fixed":          # one "subject" holding all points: pooled OLS == its own slope test
        # reshape so the kernel's pooling runs over >= 2 groups of equal size when possible
        N, J = (2, J // 2) if J % 2 == 0 else (1, J)
        x, y = x.reshape(N, J), y.reshape(N, J)
    av = torch.from_numpy(y).cuda().view(1, N, 1, J, 1)
    sc = torch.from_numpy(x).cuda().view(1, N, J, 1)
    rec = torch.zeros(1, N, 1, device="cuda")
    p, coef, betas = daa.daa_regression(av, sc, rec, reg_method=method)
    subjects_betas = None
    if method == "hierarchical":
        subjects_betas = pd.DataFrame({groups_name: groups, "beta": betas.view(-1).cpu().numpy()})
    return float(p.view(-1)[0]), float(coef.view(-1)[0]), subjects_betas
