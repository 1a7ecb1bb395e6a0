"""The CPU oracle (oracle/mopoe_oracle.py, oracle/daa_oracle.py) against the golden vectors that
oracle/make_golden.py recorded from the UNMODIFIED reference (tests/golden/).  Runs anywhere."""
import os

import numpy as np
import pytest
import torch

from oracle import cases, daa_oracle, mopoe_oracle as mo
from helpers import GOLDEN, RTOL, assert_digest_close, digest, load_golden

GOLD = load_golden()


@pytest.mark.parametrize("name", sorted(cases.ELBO_CASES))
def test_elbo_terms_and_gradients(name):
    case = cases.ELBO_CASES[name]
    want = GOLD["elbo"][name]
    spec = cases.spec_of(case)
    params = mo.init_params(spec, seed=case["seed"])
    batch, eps = cases.inputs_of(case, spec)
    out, grads, used = mo.elbo_and_grads(params, spec, batch, eps)
    assert abs(float(out["total_loss"]) - want["total_loss"]) <= RTOL * abs(want["total_loss"])
    assert abs(float(out["joint_divergence"]) - want["joint_divergence"]) <= RTOL * abs(want["joint_divergence"])
    for k, v in want["log_probs"].items():
        assert abs(float(out["log_probs"][k]) - v) <= RTOL * abs(v), k
    assert set(out["klds"]) == set(want["klds"])
    for k, v in want["klds"].items():
        assert abs(float(out["klds"][k]) - v) <= RTOL * abs(v), k
    for k, v in want["grads"].items():
        if v is None:
            assert not used[k], k
        else:
            assert used[k], k
            assert_digest_close(digest(grads[k]), v, what=k)
    assert_digest_close(digest(out["results"]["latents"]["joint"][0]), want["joint_mu"])
    for k, v in want["rec_loc"].items():
        assert_digest_close(digest(out["results"]["rec"][k][0]), v, what=k)


@pytest.mark.parametrize("name", ["hbn_joint_elbo_fact_01", "hbn_poe_fact_0", "hbn_moe_nofact_1",
                                  "stress_joint_elbo_13"])
def test_two_adam_steps(name):
    """oracle Adam == torch.optim.Adam as configured by experiment.py:268-271."""
    case = cases.ELBO_CASES[name]
    want = GOLD["elbo"][name]
    spec = cases.spec_of(case)
    params = mo.init_params(spec, seed=case["seed"])
    b1, e1 = cases.inputs_of(case, spec)
    b2, e2 = cases.inputs_of(dict(case, data_seed=case["data_seed"] + 1), spec)
    new, opt, losses = mo.train_steps(params, spec, [b1, b2], [e1, e2], lr=0.002)
    assert abs(float(losses[1]["total_loss"]) - want["loss_step2"]) <= RTOL * abs(want["loss_step2"])
    for k, v in want["params_after_2_steps"].items():
        assert_digest_close(digest(new[k]), v, what=k)


@pytest.mark.parametrize("name", sorted(cases.FORWARD_CASES))
def test_forward(name):
    case = cases.FORWARD_CASES[name]
    want = GOLD["forward"][name]
    spec = cases.spec_of(case)
    params = mo.init_params(spec, seed=case["seed"])
    batch, eps = cases.inputs_of(case, spec)
    with torch.no_grad():
        res = mo.forward(params, spec, batch, eps[0], sample_latents=case.get("sample_latents", True),
                         use_expert=case.get("use_expert"))
    assert_digest_close(digest(res["latents"]["joint"][0]), want["joint_mu"])
    assert_digest_close(digest(res["latents"]["joint"][1]), want["joint_logvar"])
    assert set(res["latents"]["subsets"]) == set(want["subsets"])
    for k, (mu, lv) in want["subsets"].items():
        assert_digest_close(digest(res["latents"]["subsets"][k][0]), mu, what=k)
        assert_digest_close(digest(res["latents"]["subsets"][k][1]), lv, what=k)
    assert abs(float(res["joint_divergence"]) - want["joint_divergence"]) <= RTOL * abs(want["joint_divergence"])
    for k, v in want["rec_loc"].items():
        assert_digest_close(digest(res["rec"][k][0]), v, what=k)
        assert_digest_close(digest(res["rec"][k][1].expand_as(res["rec"][k][0])), want["rec_scale"][k], what=k)


@pytest.mark.parametrize("name", sorted(cases.DAA_CASES))
def test_daa_avatars(name):
    case = cases.DAA_CASES[name]
    gold = np.load(os.path.join(GOLDEN, "reference_daa_%s.npz" % name))
    spec = cases.spec_of(case)
    params = mo.init_params(spec, seed=case["seed"])
    src, dst, eb, es, ea = cases.daa_inputs_of(case, spec)
    av, sc, rc = daa_oracle.daa_generate(params, spec, src, dst, eb, es, ea,
                                         sample_latents=case["sample_latents"])
    scale = np.abs(gold["avatars_sub"]).max()
    assert np.abs(av[..., ::cases.DAA_ROI_STRIDE] - gold["avatars_sub"]).max() <= RTOL * scale
    assert np.abs(sc - gold["sampled_scores"]).max() <= RTOL * np.abs(gold["sampled_scores"]).max()
    assert np.abs(rc - gold["reconstructions"]).max() <= RTOL * np.abs(gold["reconstructions"]).max()


def test_selection_bounds_known_values():
    # SURVEY.md "hard parts": measured on the reference expression
    assert mo.selection_bounds(256, 3) == [0, 85, 170, 256]
    assert mo.selection_bounds(50, 3) == [0, 16, 32, 50]
    b = mo.selection_bounds(65536, 15)
    assert [b[i + 1] - b[i] for i in range(15)] == [4369] * 14 + [4370]
