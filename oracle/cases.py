"""Seeded parity cases shared by oracle/make_golden.py (reference side, build container) and the
tests (oracle + CUDA side, anywhere).  TEST INFRASTRUCTURE.

Everything is regenerated from seeds: weights by mopoe_oracle.init_params (torch CPU generator),
data and noise by numpy default_rng -> identical tensors on both sides without shipping them.
"""
import itertools

import numpy as np
import torch

from . import mopoe_oracle as mo

HBN = dict(dims=[7, 444], style_dims=[3, 20], latent_dim=20, mod_names=["clinical", "rois"])
STRESS = dict(dims=[7, 444, 24, 148], style_dims=[3, 20, 3, 20], latent_dim=20,
              mod_names=["clinical", "rois", "modc", "modd"])


def _case(base, method, factorized, present, n_rows, seed, data_seed, **kw):
    c = dict(base, method=method, factorized=factorized, present=list(present), n_rows=n_rows,
             seed=seed, data_seed=data_seed)
    c.update(kw)
    return c


ELBO_CASES = {}
for _i, (_m, _f, _p) in enumerate(itertools.product(["joint_elbo", "moe", "poe"], [True, False],
                                                     [(0, 1), (0,), (1,)])):
    ELBO_CASES["hbn_%s_%s_%s" % (_m, "fact" if _f else "nofact", "".join(map(str, _p)))] = _case(
        HBN, _m, _f, _p, 256, 10 + _i, 100 + _i)
ELBO_CASES["hbn_joint_elbo_fact_01_tail37"] = _case(HBN, "joint_elbo", True, (0, 1), 37, 40, 140)
ELBO_CASES["hbn_joint_elbo_fact_01_fixedscale"] = _case(HBN, "joint_elbo", True, (0, 1), 64, 41, 141,
                                                        learn_output_scale=False)
for _i, (_m, _p) in enumerate(itertools.product(["joint_elbo", "moe", "poe"],
                                                [(0, 1, 2, 3), (1, 3), (0, 2, 3)])):
    ELBO_CASES["stress_%s_%s" % (_m, "".join(map(str, _p)))] = _case(STRESS, _m, True, _p, 96, 50 + _i, 150 + _i)

# method="jsd" (BaseMMVae.py:51-54,81-93,217-223, mm_div.py:23-35,69-89)
for _i, (_f, _p) in enumerate(itertools.product([True, False], [(0, 1), (0,), (1,)])):
    ELBO_CASES["hbn_jsd_%s_%s" % ("fact" if _f else "nofact", "".join(map(str, _p)))] = _case(
        HBN, "jsd", _f, _p, 256, 80 + _i, 180 + _i)
for _i, _p in enumerate([(0, 1, 2, 3), (1, 3), (0, 2, 3)]):
    ELBO_CASES["stress_jsd_%s" % "".join(map(str, _p))] = _case(STRESS, "jsd", True, _p, 96, 90 + _i, 190 + _i)

# architectures outside the train_exp defaults (networks.py:16-20,51-59, modality.py:18-30): the layered path
ARCH_CASES = {
    "enc2": dict(n_hidden_enc=2), "enc0": dict(n_hidden_enc=0), "dec1": dict(n_hidden_dec=1),
    "enc3_dec2": dict(n_hidden_enc=3, n_hidden_dec=2), "samplescale": dict(sample_scale=True),
    "samplescale_dec1": dict(sample_scale=True, n_hidden_dec=1), "laplace": dict(likelihood="laplace"),
    "enc0_dec1_samplescale_laplace": dict(n_hidden_enc=0, n_hidden_dec=1, sample_scale=True, likelihood="laplace"),
}
for _i, (_a, _kw) in enumerate(ARCH_CASES.items()):
    _m = ["joint_elbo", "poe", "moe", "jsd"][_i % 4]
    ELBO_CASES["hbn_%s_%s_01" % (_a, _m)] = _case(HBN, _m, True, (0, 1), 96, 200 + _i, 300 + _i, **_kw)
    ELBO_CASES["hbn_%s_joint_elbo_1" % _a] = _case(HBN, "joint_elbo", _i % 2 == 0, (1,), 64, 220 + _i, 320 + _i, **_kw)
ELBO_CASES["stress_enc2_dec1_samplescale_poe_023"] = _case(STRESS, "poe", True, (0, 2, 3), 48, 240, 340, n_hidden_enc=2,
                                                            n_hidden_dec=1, sample_scale=True)

FORWARD_CASES = {
    "hbn_n50_sampled": _case(HBN, "joint_elbo", True, (0, 1), 50, 60, 160),
    "hbn_n50_mean": _case(HBN, "joint_elbo", True, (0, 1), 50, 61, 161, sample_latents=False),
    "hbn_n50_expert_rois": _case(HBN, "joint_elbo", True, (0, 1), 50, 62, 162, use_expert="rois"),
    "hbn_n50_moe": _case(HBN, "moe", True, (0, 1), 50, 63, 163),
    "hbn_n50_poe_clinical_only": _case(HBN, "poe", False, (0,), 50, 64, 164),
    "stress_n33": _case(STRESS, "joint_elbo", True, (0, 1, 2, 3), 33, 65, 165),
    "hbn_n50_jsd": _case(HBN, "jsd", True, (0, 1), 50, 66, 166),
    "hbn_n50_jsd_mean": _case(HBN, "jsd", True, (0, 1), 50, 67, 167, sample_latents=False),
    "hbn_n50_enc2_dec1_samplescale": _case(HBN, "joint_elbo", True, (0, 1), 50, 68, 168, n_hidden_enc=2, n_hidden_dec=1,
                                           sample_scale=True),
}

DAA_ROI_STRIDE = 16

DAA_CASES = {
    "joint_elbo": dict(_case(HBN, "joint_elbo", True, (0, 1), 50, 70, 170), n_val=2, n_base=5,
                       n_samples=6, sample_latents=True),
    "moe_nofact": dict(_case(HBN, "moe", False, (0, 1), 50, 71, 171), n_val=1, n_base=3,
                       n_samples=4, sample_latents=True),
    "poe_mean": dict(_case(HBN, "poe", True, (0, 1), 20, 72, 172), n_val=1, n_base=3,
                     n_samples=4, sample_latents=False),
    "jsd": dict(_case(HBN, "jsd", True, (0, 1), 30, 73, 173), n_val=1, n_base=3,
                n_samples=4, sample_latents=True),
    "dec1_samplescale": dict(_case(HBN, "joint_elbo", True, (0, 1), 18, 74, 174, n_hidden_dec=1, sample_scale=True),
                             n_val=2, n_base=4, n_samples=5, sample_latents=True),
    "enc2_moe_mean": dict(_case(HBN, "moe", True, (0, 1), 16, 75, 175, n_hidden_enc=2), n_val=1, n_base=3,
                          n_samples=4, sample_latents=False),
}


def spec_kwargs(case):
    return dict(dims=case["dims"],
                style_dims=case["style_dims"] if case["factorized"] else [0] * len(case["dims"]),
                latent_dim=case["latent_dim"], method=case["method"], mod_names=case["mod_names"],
                learn_output_scale=case.get("learn_output_scale", True),
                n_hidden_enc=case.get("n_hidden_enc", 1), n_hidden_dec=case.get("n_hidden_dec", 0),
                sample_scale=case.get("sample_scale", False), likelihood=case.get("likelihood", "normal"))


def spec_of(case):
    return mo.ModelSpec(**spec_kwargs(case))


def inputs_of(case, spec):
    """-> (batch dict of present modalities, eps (n_pass, N, E)) as float32 torch tensors."""
    rng = np.random.default_rng(case["data_seed"])
    N = case["n_rows"]
    batch = {}
    for m in range(spec.n_mods):
        x = rng.standard_normal((N, spec.dims[m])).astype(np.float32)
        if m in case["present"]:
            batch[spec.mod_names[m]] = torch.from_numpy(x)
    n_pass = 1 + spec.n_mods if spec.method == "poe" else 1
    eps = torch.from_numpy(rng.standard_normal((n_pass, N, spec.eps_width)).astype(np.float32))
    return batch, eps


def daa_inputs_of(case, spec):
    rng = np.random.default_rng(case["data_seed"])
    nv, N, J, Mb = case["n_val"], case["n_rows"], case["n_samples"], case["n_base"]
    C, R, E = spec.dims[0], spec.dims[1], spec.eps_width
    f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    return f(nv, N, C), f(nv, N, R), f(nv, Mb, N, E), f(nv, J, N, C), f(nv, J, C, N, E)
