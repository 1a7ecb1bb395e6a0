"""Golden batch plans of the reference's MissingModalitySampler (multimodal_cohort/dataset.py:274-354), produced by
running the UNMODIFIED reference class in the build container under fixed numpy seeds.  TEST INFRASTRUCTURE ONLY.
Writes tests/golden/reference_epoch_plans.json; `tests/test_host_logic.py` compares `mopoe_b200.data.epoch_plan`
with it (the reference tree does not travel to the GPU box).  usage: python -m oracle.make_golden_sampler"""
import json
import os
from types import SimpleNamespace

import numpy as np

from oracle import ref_harness

CASES = [dict(n_both=700, n_clinical_only=130, n_rois_only=70, batch_size=256, seed=0),
         dict(n_both=700, n_clinical_only=130, n_rois_only=70, batch_size=64, seed=7),
         dict(n_both=300, n_clinical_only=0, n_rois_only=0, batch_size=128, seed=3),      # no missing blocks
         dict(n_both=256, n_clinical_only=10, n_rois_only=300, batch_size=256, seed=11)]  # exact multiple + tails


def has_matrix(case):
    n = case["n_both"] + case["n_clinical_only"] + case["n_rois_only"]
    has = np.ones((2, n), bool)
    has[1, case["n_both"]:case["n_both"] + case["n_clinical_only"]] = False
    has[0, case["n_both"] + case["n_clinical_only"]:] = False
    return has


def reference_plan(case):
    ref_harness.install()
    from multimodal_cohort.dataset import MissingModalitySampler
    has = has_matrix(case)
    masks = (has * np.array([1, 2])[:, None]).sum(0)
    subsets = sorted(set(masks.tolist()))
    ds = SimpleNamespace(modality_subsets=subsets,
                         idx_per_modality_subset=[np.flatnonzero(masks == m).tolist() for m in subsets])
    s = MissingModalitySampler.__new__(MissingModalitySampler)   # (Sampler.__init__ only stores the data source)
    s.dataset, s.indices, s.batch_size, s.stratify, s.discretize, s.seed = ds, None, case["batch_size"], None, None, 42
    np.random.seed(case["seed"])
    return [np.asarray(b).astype(int).tolist() for b in s.__iter__()]


def main():
    out = [dict(case=c, plan=reference_plan(c)) for c in CASES]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_epoch_plans.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, [len(o["plan"]) for o in out])


if __name__ == "__main__":
    main()
