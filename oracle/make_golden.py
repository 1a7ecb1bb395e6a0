"""Run the UNMODIFIED reference (from /root/reference, this container only) on seeded inputs with
injected noise and store what it returns as small golden fixtures under tests/golden/.

TEST INFRASTRUCTURE.  Usage:  python -m oracle.make_golden
The fixtures are what `tests/test_golden.py` checks the oracle restatement (CPU) and the CUDA
path (GPU) against on the GPU box, where /root/reference does not exist.

Inputs are regenerated from seeds at test time (oracle.cases): weights = oracle init_params(seed)
loaded into the reference model with load_state_dict, x/eps = numpy default_rng(seed).
Stored per case: every loss term, per-parameter gradient digests (sum, abs-sum, first 6 values),
and a slice of the forward outputs.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, mopoe_oracle as mo, ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ref_model_for(case):
    spec = cases.spec_of(case)
    flags = rh.make_flags(input_dims=spec.dims, latent_dim=spec.latent_dim,
                          style_dim=case["style_dims"] if case["factorized"] else [3] * len(spec.dims),
                          method=spec.method, factorized_representation=case["factorized"],
                          learn_output_scale=spec.learn_output_scale, num_hidden_layer_encoder=spec.n_hidden_enc,
                          num_hidden_layer_decoder=spec.n_hidden_dec, likelihood=spec.likelihood,
                          out_scale_per_subject=spec.sample_scale)
    model, exp = rh.build_reference_model(flags, seed=0, mod_names=spec.mod_names)
    params = mo.init_params(spec, seed=case["seed"])
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model, exp, spec, params


def digest(t):
    f = t.detach().reshape(-1).double()
    return [float(f.sum()), float(f.abs().sum())] + [float(v) for v in f[:6]]


def run_elbo_case(case):
    import run_epochs
    model, exp, spec, params = ref_model_for(case)
    batch, eps = cases.inputs_of(case, spec)
    present = [m for m, n in enumerate(spec.mod_names) if n in batch]
    with rh.InjectedNoise(mo.reference_eps_list(spec, present, eps)) as inj:
        out = run_epochs.basic_routine_epoch(exp, 0, ({k: v.clone() for k, v in batch.items()}, None, None))
        assert not inj.eps
    model.zero_grad()
    out["total_loss"].backward()
    res = out["results"]
    g = {"total_loss": float(out["total_loss"]),
         "joint_divergence": float(res["joint_divergence"]),
         "individual_divs": [float(v) for v in res["individual_divs"]],
         "log_probs": {k: float(v) for k, v in out["log_probs"].items()},
         "klds": {k: float(v) for k, v in out["klds"].items()},
         "grads": {k: (None if p.grad is None else digest(p.grad)) for k, p in model.named_parameters()},
         "joint_mu": digest(res["latents"]["joint"][0]),
         "joint_logvar": digest(res["latents"]["joint"][1]),
         "rec_loc": {k: digest(v.loc) for k, v in res["rec"].items()}}
    # two Adam steps with torch.optim.Adam as experiment.py:268-271 configures it
    opt = torch.optim.Adam(list(model.parameters()), lr=0.002, betas=(0.9, 0.999))
    opt.step()
    batch2, eps2 = cases.inputs_of(dict(case, data_seed=case["data_seed"] + 1), spec)
    with rh.InjectedNoise(mo.reference_eps_list(spec, present, eps2)):
        out2 = run_epochs.basic_routine_epoch(exp, 0, ({k: v.clone() for k, v in batch2.items()}, None, None))
    opt.zero_grad()
    out2["total_loss"].backward()
    opt.step()
    g["loss_step2"] = float(out2["total_loss"])
    g["params_after_2_steps"] = {k: digest(p) for k, p in model.named_parameters()}
    return g


def run_forward_case(case):
    model, exp, spec, params = ref_model_for(case)
    batch, eps = cases.inputs_of(case, spec)
    present = [m for m, n in enumerate(spec.mod_names) if n in batch]
    model.eval()
    with torch.no_grad(), rh.InjectedNoise(mo.reference_eps_list(spec, present, eps, True)
                                           if case.get("sample_latents", True) else []):
        res = model(batch, sample_latents=case.get("sample_latents", True),
                    use_expert=case.get("use_expert"))
    return {"joint_mu": digest(res["latents"]["joint"][0]),
            "joint_logvar": digest(res["latents"]["joint"][1]),
            "subsets": {k: [digest(v[0]), digest(v[1])] for k, v in res["latents"]["subsets"].items()},
            "joint_divergence": float(res["joint_divergence"]),
            "rec_loc": {k: digest(v.loc) for k, v in res["rec"].items()},
            "rec_scale": {k: digest(v.scale) for k, v in res["rec"].items()}}


def run_daa_case(case):
    """workflow.py:361-427 driven with the real reference model and injected noise."""
    model, exp, spec, params = ref_model_for(case)
    src, dst, eb, es, ea = cases.daa_inputs_of(case, spec)
    model.eval()
    n_val, N, C = src.shape
    J, Mb, R = es.shape[1], eb.shape[1], dst.shape[2]
    avatars = np.zeros((n_val, N, C, J, R), np.float32)
    sampled = np.zeros((n_val, N, J, C), np.float32)
    recons = np.zeros((n_val, N, R), np.float32)
    present = [0, 1]
    with torch.no_grad():
        for v in range(n_val):
            data = {"clinical": src[v], "rois": dst[v]}
            cl, cs, rl = [], [], []
            for p in range(Mb):
                with rh.InjectedNoise(mo.reference_eps_list(spec, present, eb[v, p][None], True)):
                    rec = model(data, sample_latents=True)["rec"]
                cl.append(rec["clinical"].loc.unsqueeze(0))
                cs.append(rec["clinical"].scale.unsqueeze(0))
                rl.append(rec["rois"].loc.unsqueeze(0))
            loc_hat = torch.cat(cl).mean(0)
            scale_hat = torch.cat(cs).mean(0)
            recons[v] = torch.cat(rl).mean(0).numpy()
            scores = loc_hat + scale_hat * es[v]          # Normal(loc,scale).sample with injected eps
            for j in range(J):
                for c in range(C):
                    cdata = data["clinical"].clone()
                    cdata[:, c] = scores[j, :, c]
                    noise = (mo.reference_eps_list(spec, present, ea[v, j, c][None], True)
                             if case["sample_latents"] else [])
                    with rh.InjectedNoise(noise):
                        rec = model({"clinical": cdata, "rois": data["rois"]},
                                    sample_latents=case["sample_latents"])["rec"]
                    avatars[v, :, c, j] = rec["rois"].loc.numpy()
            sampled[v] = np.swapaxes(scores.numpy(), 0, 1)
    return avatars, sampled, recons


def main():
    os.makedirs(OUT, exist_ok=True)
    rh.install()
    gold = {"elbo": {}, "forward": {}}
    for name, case in cases.ELBO_CASES.items():
        gold["elbo"][name] = run_elbo_case(case)
        print("elbo", name, gold["elbo"][name]["total_loss"])
    for name, case in cases.FORWARD_CASES.items():
        gold["forward"][name] = run_forward_case(case)
        print("forward", name)
    with open(os.path.join(OUT, "reference_elbo_forward.json"), "w") as f:
        json.dump(gold, f, indent=0, sort_keys=True)
    for name, case in cases.DAA_CASES.items():
        av, sc, rc = run_daa_case(case)
        # keep the fixture small: every 16th ROI column of the avatar tensor
        np.savez_compressed(os.path.join(OUT, "reference_daa_%s.npz" % name),
                            avatars_sub=av[..., ::cases.DAA_ROI_STRIDE].astype(np.float32),
                            sampled_scores=sc, reconstructions=rc)
        print("daa", name, av.shape, float(np.abs(av).mean()))


if __name__ == "__main__":
    main()
