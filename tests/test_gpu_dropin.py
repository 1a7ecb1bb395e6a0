"""The reference-interface mirrors on the GPU: VAE drop-in, basic_routine_epoch, make_regression,
train_exp / daa_exp end to end on a small synthetic cohort."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import cases, daa_oracle, mopoe_oracle as mo
from helpers import RTOL, load_golden

pytestmark = pytest.mark.gpu


def _flags(method="joint_elbo", factorized=True):
    f = SimpleNamespace(input_dim=[7, 444], style_dim=[3, 20], class_dim=20, factorized_representation=factorized,
                        modality_poe=method == "poe", modality_moe=method == "moe", modality_jsd=method == "jsd",
                        joint_elbo=method == "joint_elbo", learn_output_scale=True, learn_output_sample_scale=False,
                        beta=1.0, beta_style=1.0, beta_content=1.0, num_hidden_layer_encoder=1,
                        num_hidden_layer_decoder=0, likelihood="normal", initial_out_logvar=-3.0, dropout_rate=0.0,
                        num_models=1, dir_checkpoints="")
    return f


def _model(method="joint_elbo", seed=11):
    from mopoe_b200.model import VAE
    flags = _flags(method)
    mods = {"clinical": SimpleNamespace(name="clinical"), "rois": SimpleNamespace(name="rois")}
    model = VAE(flags, mods).cuda()
    ospec = mo.ModelSpec(method=method)
    params = mo.init_params(ospec, seed=seed)
    model.load_state_dict(params, strict=True)
    return model, ospec, params


def test_state_dict_keys_match_reference():
    model, ospec, params = _model()
    assert list(model.state_dict().keys()) == list(mo.param_shapes(ospec).keys())
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == mo.param_shapes(ospec)
    # same construction order as the reference => same default init under the same torch seed
    gold = load_golden()
    assert sum(p.numel() for p in model.parameters()) == 167173


def test_forward_results_structure_and_values():
    model, ospec, params = _model()
    g = torch.Generator().manual_seed(3)
    batch = {"clinical": torch.randn(50, 7, generator=g), "rois": torch.randn(50, 444, generator=g)}
    eps = torch.randn(50, ospec.eps_width, generator=g)
    model.inject_noise(eps)
    res = model({k: v.cuda() for k, v in batch.items()}, sample_latents=True)
    with torch.no_grad():
        want = mo.forward(params, ospec, batch, eps)
    assert set(res) >= {"latents", "group_distr", "joint_divergence", "individual_divs", "dyn_prior", "rec"}
    assert set(res["latents"]) == {"modalities", "mus", "logvars", "weights", "joint", "subsets"}
    assert list(res["latents"]["subsets"]) == ["clinical", "rois", "clinical_rois"]
    assert res["latents"]["mus"].shape == (3, 50, 20)
    for k in ("clinical", "rois"):
        assert isinstance(res["rec"][k], torch.distributions.Normal)
        assert torch.allclose(res["rec"][k].loc.cpu(), want["rec"][k][0], rtol=0, atol=RTOL * float(want["rec"][k][0].abs().max()))
        assert torch.allclose(res["rec"][k].scale.cpu()[0], want["rec"][k][1][0], rtol=1e-6)
    assert torch.allclose(res["latents"]["mus"].cpu(), want["latents"]["mus"], atol=RTOL * float(want["latents"]["mus"].abs().max()))
    assert abs(float(res["joint_divergence"]) - float(want["joint_divergence"])) <= RTOL * abs(float(want["joint_divergence"]))
    # missing block: absent modality is neither encoded nor decoded
    res1 = model({"rois": batch["rois"].cuda()})
    assert list(res1["rec"]) == ["rois"] and res1["latents"]["modalities"]["clinical"] == [None, None]
    with pytest.raises(Exception):
        model(batch)          # CPU tensors: no fallback


def test_forward_results_of_the_jsd_model():
    """method="jsd" through the drop-in model: mixture = unimodal experts + the prior, individual_divs and dyn_prior
    as divergence_dynamic_prior returns them (BaseMMVae.py:81-93,217-223)."""
    model, ospec, params = _model("jsd")
    g = torch.Generator().manual_seed(4)
    batch = {"clinical": torch.randn(50, 7, generator=g), "rois": torch.randn(50, 444, generator=g)}
    eps = torch.randn(50, ospec.eps_width, generator=g)
    model.inject_noise(eps)
    res = model({k: v.cuda() for k, v in batch.items()}, sample_latents=True)
    with torch.no_grad():
        want = mo.forward(params, ospec, batch, eps)
    assert res["latents"]["mus"].shape == (3, 50, 20) and float(res["latents"]["mus"][2].abs().max()) == 0.0
    assert torch.allclose(res["latents"]["weights"].cpu(), want["latents"]["weights"])
    tol = lambda t: RTOL * float(t.abs().max())
    assert torch.allclose(res["latents"]["joint"][0].cpu(), want["latents"]["joint"][0], rtol=0, atol=tol(want["latents"]["joint"][0]))
    assert float(res["latents"]["joint"][0][32:].abs().max()) == 0.0          # rows [32, 50) sample from the prior
    assert torch.allclose(res["individual_divs"].cpu(), want["individual_divs"], rtol=RTOL)
    assert abs(float(res["joint_divergence"]) - float(want["joint_divergence"])) <= RTOL * abs(float(want["joint_divergence"]))
    for a, w in zip(res["dyn_prior"], want["dyn_prior"]):
        assert torch.allclose(a.cpu(), w, rtol=0, atol=tol(w))
    assert torch.allclose(res["rec"]["rois"].loc.cpu(), want["rec"]["rois"][0], rtol=0, atol=tol(want["rec"]["rois"][0]))


def test_philox_seed_draws_fresh_noise_per_call():
    """In-kernel generator through the drop-in interface: every call draws new noise (the reference draws from
    the advancing global generator, BaseMMVae.py:37-40), reproducibly for a given philox_seed; seed 0 is a seed."""
    from mopoe_b200 import run_epochs
    model, ospec, params = _model()
    g = torch.Generator().manual_seed(3)
    batch = {"clinical": torch.randn(50, 7, generator=g).cuda(), "rois": torch.randn(50, 444, generator=g).cuda()}
    for seed in (0, 77):
        model.philox_seed, model._philox_calls = seed, 0
        a = model(batch)["class_embeddings"].clone()
        b = model(batch)["class_embeddings"].clone()
        assert not torch.equal(a, b)
        model._philox_calls = 0
        assert torch.equal(model(batch)["class_embeddings"], a)
        exp = SimpleNamespace(models=model, flags=model.flags)
        with torch.no_grad():
            l0 = float(run_epochs.basic_routine_epoch(exp, 0, (dict(batch), None, None))["total_loss"])
            l1 = float(run_epochs.basic_routine_epoch(exp, 0, (dict(batch), None, None))["total_loss"])
        assert l0 != l1
    model.philox_seed, model._philox_calls = 0, 0
    z0 = model(batch)["class_embeddings"].clone()
    model.philox_seed, model._philox_calls = 77, 0
    assert not torch.equal(model(batch)["class_embeddings"], z0)


@pytest.mark.parametrize("method", ["joint_elbo", "poe", "moe", "jsd"])
def test_basic_routine_epoch_backward_and_torch_adam(method):
    from mopoe_b200 import run_epochs
    model, ospec, params = _model(method)
    exp = SimpleNamespace(models=model, flags=model.flags)
    g = torch.Generator().manual_seed(5)
    batch = {"clinical": torch.randn(96, 7, generator=g), "rois": torch.randn(96, 444, generator=g)}
    eps = torch.randn(3 if method == "poe" else 1, 96, ospec.eps_width, generator=g)
    model.inject_noise(eps)
    opt = torch.optim.Adam(model.parameters(), lr=0.002, betas=(0.9, 0.999))
    out = run_epochs.basic_routine_epoch(exp, 0, ({k: v.clone() for k, v in batch.items()}, None, None))
    assert set(out) == {"results", "log_probs", "total_loss", "klds"}
    opt.zero_grad()
    out["total_loss"].backward()
    want, grads, used = mo.elbo_and_grads(params, ospec, batch, eps)
    assert abs(float(out["total_loss"]) - float(want["total_loss"])) <= RTOL * abs(float(want["total_loss"]))
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        err = float((p.grad.cpu() - grads[k]).abs().max() / (grads[k].abs().max() + 1e-30))
        assert err <= RTOL, (k, err)
    for k, v in want["klds"].items():
        assert abs(float(out["klds"][k]) - float(v)) <= RTOL * abs(float(v))
    before = model.state_dict()["encoders.rois.class_mu.weight"].clone()
    opt.step()                                  # stock torch optimiser works on the flat-buffer views
    assert not torch.equal(before, model.state_dict()["encoders.rois.class_mu.weight"])
    # missing block: gradients of the absent modality stay None (torch.optim.Adam then skips them)
    opt.zero_grad(set_to_none=True)
    out = run_epochs.basic_routine_epoch(exp, 0, ({"rois": batch["rois"].clone()}, None, None))
    out["total_loss"].backward()
    assert model.encoders["clinical"].class_mu.weight.grad is None
    assert model.encoders["rois"].class_mu.weight.grad is not None


def test_make_regression_mirror():
    import pandas as pd
    from mopoe_b200 import stat_utils
    rng = np.random.default_rng(0)
    G, J = 9, 40
    x = rng.standard_normal((G, J)).astype(np.float32)
    y = (0.3 * x + 0.1 * rng.standard_normal((G, J)) + rng.standard_normal((G, 1))).astype(np.float32)
    df = pd.DataFrame({"participant_id": np.repeat(np.arange(G), J), "sampled_score": x.reshape(-1), "roi_avatar": y.reshape(-1)})
    p, coef, betas = stat_utils.make_regression(df, "sampled_score", "roi_avatar", groups_name="participant_id", method="hierarchical")
    pw, cw, bw = daa_oracle.hierarchical_regression(y[None, :, None, :, None], x[None, :, :, None])
    assert abs(coef - cw[0, 0, 0]) <= 1e-9 and abs(np.log(p) - np.log(pw[0, 0, 0])) <= 1e-7
    assert np.allclose(betas["beta"].to_numpy(), bw[0, 0, :, 0], rtol=1e-9)
    p, coef, _ = stat_utils.make_regression(df, "sampled_score", "roi_avatar", method="fixed")
    from scipy import stats as sps
    lr = sps.linregress(x.reshape(-1).astype(np.float64), y.reshape(-1).astype(np.float64))
    assert abs(coef - lr.slope) <= 1e-9 and abs(np.log(p) - np.log(lr.pvalue)) <= 1e-6
    with pytest.raises(NotImplementedError):
        stat_utils.make_regression(df, "sampled_score", "roi_avatar", groups_name="participant_id", method="mixed")


def test_train_exp_then_daa_exp_end_to_end(tmp_path):
    from mopoe_b200 import data, workflow, daa
    ds, out = str(tmp_path / "data"), str(tmp_path / "out")
    os.makedirs(out)
    data.write_dataset(ds, data.make_cohort(n_both=640, n_clinical_only=128, n_rois_only=64, standardize=False))
    run = workflow.train_exp("hbn", ds, out, [7, 444], num_epochs=6, batch_size=128, method="joint_elbo", data_seed=3)
    rundir = os.path.join(out, run)
    assert os.path.isfile(os.path.join(rundir, "flags.rar"))
    assert os.path.isfile(os.path.join(rundir, "checkpoints", "0004", "model"))
    assert os.path.isfile(os.path.join(rundir, "checkpoints", "0005", "model"))
    assert os.path.isfile(os.path.join(rundir, "checkpoints", "enc_rois"))
    sd = torch.load(os.path.join(rundir, "checkpoints", "0005", "model"))
    assert list(sd.keys()) == list(mo.param_shapes(mo.ModelSpec()).keys())
    tr = np.load(os.path.join(rundir, "logs", "scalars_train_model0.npy"))
    assert np.isfinite(tr).all() and tr[-1, 0] < tr[0, 0]          # the loss goes down
    resdir = workflow.daa_exp("hbn", ds, out, run, n_validation=3, n_samples=12, n_subjects=20, M=30, trust_level=0.7)
    av = np.load(os.path.join(resdir, "rois_digital_avatars.npy"))
    sc = np.load(os.path.join(resdir, "sampled_scores.npy"))
    assert av.shape == (3, 20, 7, 12, 444) and av.dtype == np.float32 and sc.shape == (3, 20, 12, 7)
    assert np.load(os.path.join(resdir, "rois_reconstructions.npy")).shape == (3, 20, 444)
    p = np.load(os.path.join(resdir, "pvalues.npy")); c = np.load(os.path.join(resdir, "coefs.npy"))
    assert p.shape == c.shape == (3, 7, 444) and p.dtype == np.float64
    pw, cw, _ = daa_oracle.hierarchical_regression(av, sc)          # statistics recomputed from the files
    assert np.allclose(c, cw, rtol=1e-9, atol=1e-12)
    assert np.all(np.abs(np.log(p) - np.log(pw)) <= 1e-7 * np.maximum(1.0, np.abs(np.log(pw))))
    allc = np.load(os.path.join(resdir, "all_coefs.npy"), allow_pickle=True)
    assert allc.shape[:2] == (3, 7)
    sig = open(os.path.join(resdir, "significant_rois.tsv")).read().splitlines()
    assert sig[0].split("\t") == ["metric", "roi", "score"]
    want = daa_oracle.significant(pw, 0.7)
    assert len(sig) - 1 == int(want.sum())
    meta = np.load(os.path.join(resdir, "metadatas.npy"), allow_pickle=True)
    assert meta.shape[:2] == (3, 20)
    # the files through the reference's own reader expressions (workflow.py:611-630 anova_exp, analyze_avatars.py:59-78)
    import pandas as pd
    rois_names = np.load(os.path.join(ds, "rois_names.npy"), allow_pickle=True)
    clinical_names = np.load(os.path.join(ds, "clinical_names.npy"), allow_pickle=True)
    modified = [n.replace("&", "_").replace("-", "_") for n in rois_names]
    all_coefs = np.load(os.path.join(resdir, "all_coefs.npy"), allow_pickle=True)[np.newaxis]
    idx_sign = ((p < 0.05 / len(rois_names) / len(clinical_names)).sum(axis=0) >= 3 * 0.7)
    assert np.array_equal(idx_sign, want)
    for val_idx in range(3):
        for score_idx in range(len(clinical_names)):
            coefs_df = pd.DataFrame(all_coefs[0][val_idx][score_idx], columns=["participant_id", "site"] + modified)
            coefs_df[modified] = coefs_df[modified].astype(float)
            assert coefs_df.shape == (20, 2 + 444)
            assert np.allclose(coefs_df[modified].to_numpy().mean(0), c[val_idx, score_idx], rtol=1e-9, atol=1e-12)
            assert list(coefs_df["participant_id"]) == list(meta[val_idx][:, 0])
    da = np.load(os.path.join(resdir, "rois_digital_avatars.npy"), mmap_mode="r")[1]
    scores, metadata = sc[1], meta[1]
    assert da.shape == (20, 7, 12, 444) and scores[5].shape == (12, 7) and metadata.shape[0] == 20
    # sampling_strategy "linear" (workflow.py:337-346): the same ramp between train-set quantiles for every subject
    lin = workflow.daa_exp("hbn", ds, out, run, sampling_strategy="linear", n_validation=2, n_samples=10, n_subjects=20, M=30)
    assert "sampling_linear" in lin
    sl = np.load(os.path.join(lin, "sampled_scores.npy"))
    assert sl.shape == (2, 20, 10, 7) and np.array_equal(sl[0, 0], sl[1, 7]) and np.all(np.diff(sl[0, 0], axis=0) > 0)
    avl = np.load(os.path.join(lin, "rois_digital_avatars.npy"))
    pl_, cl_, _ = daa_oracle.hierarchical_regression(avl, sl)
    assert np.allclose(np.load(os.path.join(lin, "coefs.npy")), cl_, rtol=1e-9, atol=1e-12)
    # rsa_exp (workflow.py:656-820) on the same run: files, shapes, and the statistics recomputed with scipy
    from oracle import rsa_oracle as ro
    rsadir = workflow.rsa_exp("hbn", ds, out, run, n_validation=2, n_subjects=40, seed=5)
    kt = np.load(os.path.join(rsadir, "kendalltau_stats.npy"))
    ld = np.load(os.path.join(rsadir, "latent_dissimilarity.npy"))
    sd_ = np.load(os.path.join(rsadir, "scores_dissimilarity.npy"))
    assert kt.shape == (1, 4, 2, 10, 2) and ld.shape == (1, 8, 40, 40) and sd_.shape == (1, 8, 10, 40, 40)
    for k in range(8):                       # entry k = (validation k // 4, latent k % 4)
        for r in range(10):
            tau, pval = ro.fit_rsa(ld[0, k], sd_[0, k, r])
            assert abs(kt[0, k % 4, k // 4, r, 0] - tau) <= 1e-13 and abs(kt[0, k % 4, k // 4, r, 1] - pval) <= 1e-9 * max(pval, 1e-300)
    tsv = pd.read_table(os.path.join(rsadir, "kendalltau_joint.tsv"))
    assert list(tsv.columns) == ["score", "pval", "pval_std", "r", "r_std"] and len(tsv) == 10
    assert list(tsv["score"][-3:]) == ["age", "sex", "site"]


def test_ensemble_train_and_daa_with_vote_prop(tmp_path):
    """num_models > 1 (workflow.py:41,203-207,290-297,524-525): k-fold ensemble, one model / scaler / Adam state per
    fold, checkpoints under model_<i>/, DAA outputs with a leading model axis, vote over models with vote_prop."""
    from mopoe_b200 import data, workflow
    ds, out = str(tmp_path / "data"), str(tmp_path / "out")
    os.makedirs(out)
    data.write_dataset(ds, data.make_cohort(n_both=480, n_clinical_only=64, n_rois_only=32, standardize=False))
    run = workflow.train_exp("hbn", ds, out, [7, 444], num_models=2, num_epochs=5, batch_size=128, data_seed=5)
    rundir = os.path.join(out, run)
    for i in range(2):
        assert os.path.isfile(os.path.join(rundir, "checkpoints", "model_%d" % i, "0004", "model"))
        tr = np.load(os.path.join(rundir, "logs", "scalars_train_model%d.npy" % i))
        assert np.isfinite(tr).all() and tr[-1, 0] < tr[0, 0]
    sd0 = torch.load(os.path.join(rundir, "checkpoints", "model_0", "0004", "model"))
    sd1 = torch.load(os.path.join(rundir, "checkpoints", "model_1", "0004", "model"))
    assert not torch.equal(sd0["encoders.rois.class_mu.weight"], sd1["encoders.rois.class_mu.weight"])   # different folds
    resdir = workflow.daa_exp("hbn", ds, out, run, n_validation=2, n_samples=8, n_subjects=16, M=20, trust_level=0.5, vote_prop=0.5)
    av = np.load(os.path.join(resdir, "rois_digital_avatars.npy"))
    p = np.load(os.path.join(resdir, "pvalues.npy"))
    assert av.shape == (2, 2, 16, 7, 8, 444) and p.shape == (2, 2, 7, 444)
    assert np.load(os.path.join(resdir, "sampled_scores.npy")).shape == (2, 2, 16, 8, 7)
    assert np.load(os.path.join(resdir, "metadatas.npy"), allow_pickle=True).shape[:3] == (2, 2, 16)
    for i in range(2):                                   # statistics recomputed from the files, per model
        pw, cw, _ = daa_oracle.hierarchical_regression(av[i], np.load(os.path.join(resdir, "sampled_scores.npy"))[i])
        assert np.all(np.abs(np.log(p[i]) - np.log(pw)) <= 1e-7 * np.maximum(1.0, np.abs(np.log(pw))))
    want = workflow.significant_votes(p, 0.5, 2, 0.5)
    sig = open(os.path.join(resdir, "significant_rois.tsv")).read().splitlines()
    assert len(sig) - 1 == int(want.sum())


_TWO_RANKS = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import mopoe_b200
from mopoe_b200 import workflow
rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ds, out, run, J = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
workflow.daa_exp("hbn", ds, out, run, n_validation=3, n_samples=J, n_subjects=20, M=30, trust_level=0.7)
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("J", [130, 12])
def test_daa_exp_two_ranks_write_the_same_files_as_one(tmp_path, J):
    """SURVEY.md 8e through the workflow: two ranks split the 21 (validation, score) units 11 / 10 (validation 1 is
    shared), each writes its slices of the avatar memmap, rank 0 the tables: every file equals the one-rank run
    (J = 130: pipelined tcgen05 kernel with unit ranges; J = 12: CUDA-core kernel)."""
    import shutil, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mopoe_b200 import data, workflow
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ds, out1, out2 = str(tmp_path / "data"), str(tmp_path / "out1"), str(tmp_path / "out2")
    os.makedirs(out1)
    data.write_dataset(ds, data.make_cohort(n_both=640, n_clinical_only=128, n_rois_only=64, standardize=False))
    run = workflow.train_exp("hbn", ds, out1, [7, 444], num_epochs=3, batch_size=128, method="joint_elbo", data_seed=3)
    shutil.copytree(out1, out2)
    res1 = workflow.daa_exp("hbn", ds, out1, run, n_validation=3, n_samples=J, n_subjects=20, M=30, trust_level=0.7)
    script = tmp_path / "two.py"
    script.write_text(_TWO_RANKS % root)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29537", str(script), ds, out2, run, str(J)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    res2 = res1.replace(out1, out2)
    for f in ("rois_digital_avatars.npy", "sampled_scores.npy", "rois_reconstructions.npy", "pvalues.npy", "coefs.npy"):
        a, b = np.load(os.path.join(res1, f)), np.load(os.path.join(res2, f))
        assert a.shape == b.shape and np.array_equal(a, b), f
    assert open(os.path.join(res1, "significant_rois.tsv")).read() == open(os.path.join(res2, "significant_rois.tsv")).read()
    a = np.load(os.path.join(res1, "all_coefs.npy"), allow_pickle=True)
    b = np.load(os.path.join(res2, "all_coefs.npy"), allow_pickle=True)
    assert a.shape == b.shape and bool((a == b).all())


def test_layered_architecture_through_the_drop_in_model_and_workflow(tmp_path):
    """num_hidden_layer_decoder=1 + out_scale_per_subject (networks.py:51-59) end to end: state-dict names of the
    reference, per-sample scales in results['rec'], basic_routine_epoch gradients vs the oracle, then train_exp -> daa_exp
    (layered sweep: hidden decoder layers make the decoder non-affine) with the statistics recomputed from the files."""
    from mopoe_b200 import data, run_epochs, workflow
    from mopoe_b200.model import VAE
    flags = _flags("joint_elbo")
    flags.num_hidden_layer_decoder, flags.learn_output_sample_scale = 1, True
    mods = {"clinical": SimpleNamespace(name="clinical"), "rois": SimpleNamespace(name="rois")}
    model = VAE(flags, mods).cuda()
    ospec = mo.ModelSpec(method="joint_elbo", n_hidden_dec=1, sample_scale=True)
    params = mo.init_params(ospec, seed=21)
    assert list(model.state_dict().keys()) == list(params.keys())
    model.load_state_dict(params, strict=True)
    g = torch.Generator().manual_seed(6)
    batch = {"clinical": torch.randn(80, 7, generator=g), "rois": torch.randn(80, 444, generator=g)}
    eps = torch.randn(1, 80, ospec.eps_width, generator=g)
    model.inject_noise(eps[0])
    res = model({k: v.cuda() for k, v in batch.items()})
    with torch.no_grad():
        want = mo.forward(params, ospec, batch, eps[0])
    assert res["rec"]["rois"].scale.shape == (80, 444)
    assert torch.allclose(res["rec"]["rois"].scale.cpu(), want["rec"]["rois"][1], rtol=1e-4)
    assert torch.allclose(res["rec"]["clinical"].loc.cpu(), want["rec"]["clinical"][0], rtol=0, atol=RTOL * float(want["rec"]["clinical"][0].abs().max()))
    exp = SimpleNamespace(models=model, flags=model.flags)
    model.inject_noise(eps)
    out = run_epochs.basic_routine_epoch(exp, 0, ({k: v.clone() for k, v in batch.items()}, None, None))
    out["total_loss"].backward()
    w, grads, used = mo.elbo_and_grads(params, ospec, batch, eps)
    assert abs(float(out["total_loss"]) - float(w["total_loss"])) <= RTOL * abs(float(w["total_loss"]))
    for k, p in model.named_parameters():
        err = float((p.grad.cpu() - grads[k]).abs().max() / (grads[k].abs().max() + 1e-30))
        assert err <= RTOL, (k, err)
    # workflow
    ds, out_dir = str(tmp_path / "data"), str(tmp_path / "out")
    os.makedirs(out_dir)
    data.write_dataset(ds, data.make_cohort(n_both=480, n_clinical_only=64, n_rois_only=32, standardize=False))
    run = workflow.train_exp("hbn", ds, out_dir, [7, 444], num_epochs=5, batch_size=128, num_hidden_layer_decoder=1,
                             out_scale_per_subject=True, data_seed=4)
    tr = np.load(os.path.join(out_dir, run, "logs", "scalars_train_model0.npy"))
    assert np.isfinite(tr).all() and tr[-1, 0] < tr[0, 0]
    sd = torch.load(os.path.join(out_dir, run, "checkpoints", "0004", "model"))
    assert list(sd.keys()) == list(params.keys())
    resdir = workflow.daa_exp("hbn", ds, out_dir, run, n_validation=2, n_samples=8, n_subjects=16, M=12, trust_level=0.5)
    av = np.load(os.path.join(resdir, "rois_digital_avatars.npy"))
    sc = np.load(os.path.join(resdir, "sampled_scores.npy"))
    assert av.shape == (2, 16, 7, 8, 444) and np.isfinite(av).all()
    pw, cw, _ = daa_oracle.hierarchical_regression(av, sc)
    assert np.allclose(np.load(os.path.join(resdir, "coefs.npy")), cw, rtol=1e-9, atol=1e-12)
