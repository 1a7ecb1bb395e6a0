"""Oracle restatement against the UNMODIFIED reference modules, executed live.  Only possible in
the build container (/root/reference is not shipped to the GPU box) -> skipped elsewhere; the
committed golden vectors (test_oracle_golden.py) carry the same evidence everywhere."""
import pytest
import torch

from oracle import mopoe_oracle as mo, ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


@pytest.mark.parametrize("method", ["joint_elbo", "moe", "poe"])
@pytest.mark.parametrize("present", [(0, 1), (0,), (1,)])
def test_basic_routine_epoch_live(method, present):
    rh.install()
    import run_epochs
    flags = rh.make_flags(method=method)
    model, exp = rh.build_reference_model(flags, seed=3)
    spec = mo.ModelSpec(dims=flags.input_dim, style_dims=flags.style_dim, method=method)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    N = 64
    x = {spec.mod_names[m]: torch.randn(N, spec.dims[m], generator=g) for m in present}
    eps = torch.randn(1 + spec.n_mods if method == "poe" else 1, N, spec.eps_width, generator=g)
    with rh.InjectedNoise(mo.reference_eps_list(spec, list(present), eps)):
        out = run_epochs.basic_routine_epoch(exp, 0, ({k: v.clone() for k, v in x.items()}, None, None))
    model.zero_grad()
    out["total_loss"].backward()
    o, gr, used = mo.elbo_and_grads(params, spec, x, eps)
    assert abs(float(o["total_loss"]) - float(out["total_loss"])) <= 1e-5 * abs(float(out["total_loss"]))
    for k, p in model.named_parameters():
        if p.grad is None:
            assert not used[k]
        else:
            assert (gr[k] - p.grad).abs().max() <= 1e-5 * p.grad.abs().max() + 1e-9, k
