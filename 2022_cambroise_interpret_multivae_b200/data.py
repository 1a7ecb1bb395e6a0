"""Synthetic HBN-shaped cohort (the reference ships no data: README.md:37-38).

numpy default_rng(42): latent u ~ N(0, I_8); clinical = u A_c + 0.5 eps (7 scores),
rois = u A_r + 0.7 eps (444 ROIs = 74 Destrieux labels x 2 hemispheres x 3 metrics) with a sparse
A_r (~15 % non-zero) so a non-trivial set of ROI-score associations exists.  2 560 subjects with both
blocks (2 048 train / 512 test), 512 clinical-only and 256 rois-only train subjects."""
import os

import numpy as np

SCORE_NAMES = ["SRS_Total", "CBCL_AB", "CBCL_AP", "CBCL_WD", "SDQ_ext", "SDQ_int", "ARI_P"]


def roi_names(n_rois=444):
    names = []
    metrics = ["thickness", "area", "meancurv"]
    k = 0
    while len(names) < n_rois:
        for hemi in ("lh", "rh"):
            for met in metrics:
                if len(names) < n_rois:
                    names.append("label%02d_%s_%s" % (k, hemi, met))
        k += 1
    return np.array(names, dtype=object)


def make_cohort(n_both=2560, n_clinical_only=512, n_rois_only=256, n_scores=7, n_rois=444, seed=42,
                standardize=True):
    """-> dict(clinical (n, 7) f32, rois (n, 444) f32, has_clinical (n,), has_rois (n,), subjects)."""
    rng = np.random.default_rng(seed)
    n = n_both + n_clinical_only + n_rois_only
    u = rng.standard_normal((n, 8))
    A_c = rng.standard_normal((8, n_scores))
    A_r = rng.standard_normal((8, n_rois)) * (rng.random((8, n_rois)) < 0.15)
    clinical = u @ A_c + 0.5 * rng.standard_normal((n, n_scores))
    rois = u @ A_r + 0.7 * rng.standard_normal((n, n_rois))
    has_c = np.ones(n, bool)
    has_r = np.ones(n, bool)
    has_r[n_both:n_both + n_clinical_only] = False
    has_c[n_both + n_clinical_only:] = False
    if standardize:   # what the loader's StandardScaler does (experiment.py:146-166)
        clinical = (clinical - clinical[has_c].mean(0)) / clinical[has_c].std(0)
        rois = (rois - rois[has_r].mean(0)) / rois[has_r].std(0)
    return dict(clinical=clinical.astype(np.float32), rois=rois.astype(np.float32), has_clinical=has_c,
                has_rois=has_r, subjects=np.array(["sub-%05d" % i for i in range(n)], dtype=object),
                n_both=n_both)


def write_dataset(datasetdir, cohort=None, seed=42):
    """Write the files the reference loaders expect (dataset.py:57-58, workflow.py:241-244):
    {clinical,rois}_data.npy, _subjects.npy, _names.npy and metadata.tsv."""
    import pandas as pd
    os.makedirs(datasetdir, exist_ok=True)
    c = cohort or make_cohort(seed=seed, standardize=False)
    for mod, has in (("clinical", c["has_clinical"]), ("rois", c["has_rois"])):
        np.save(os.path.join(datasetdir, mod + "_data.npy"), c[mod][has])
        np.save(os.path.join(datasetdir, mod + "_subjects.npy"), c["subjects"][has])
    np.save(os.path.join(datasetdir, "clinical_names.npy"), np.array(SCORE_NAMES[: c["clinical"].shape[1]], dtype=object))
    np.save(os.path.join(datasetdir, "rois_names.npy"), roi_names(c["rois"].shape[1]))
    rng = np.random.default_rng(seed + 1)
    n = len(c["subjects"])
    meta = pd.DataFrame(dict(participant_id=c["subjects"], sex=rng.integers(0, 2, n),
                             age=rng.uniform(5, 21, n).round(2), site=rng.integers(0, 4, n)))
    meta.to_csv(os.path.join(datasetdir, "metadata.tsv"), sep="\t", index=False)
    return c


def epoch_plan(has, batch_size, rng=None):
    """Batch plan of one epoch with the MissingModalitySampler contract (dataset.py:295-354): every
    batch is homogeneous in its set of present modalities; complete batches of all subsets come
    first in random order, incomplete tail batches after, drawn with numpy's global-style
    `choice(..., replace=False)` calls in the reference's order.
    has: (n_mods, n_subjects) bool.  -> list of (present_mask, row index array)."""
    rng = rng or np.random
    has = np.asarray(has, bool)
    masks = (has * (1 << np.arange(has.shape[0]))[:, None]).sum(0)
    complete, incomplete = [], []
    for mask in sorted(set(masks.tolist()) - {0}):
        pool = np.flatnonzero(masks == mask).tolist()
        while pool:
            size = min(len(pool), batch_size)
            pick = rng.choice(pool, size=size, replace=False)
            chosen = set(pick.tolist())
            pool = [i for i in pool if i not in chosen]
            (complete if size == batch_size else incomplete).append((mask, np.asarray(pick, np.int32)))
    order_c = rng.choice(len(complete), size=len(complete), replace=False) if complete else []
    order_i = rng.choice(len(incomplete), size=len(incomplete), replace=False) if incomplete else []
    return [complete[i] for i in order_c] + [incomplete[i] for i in order_i]
