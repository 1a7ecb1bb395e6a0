"""Oracle restatement against the UNMODIFIED reference modules, executed live.  Only possible in
the build container (/root/reference is not shipped to the GPU box) -> skipped elsewhere; the
committed golden vectors (test_oracle_golden.py) carry the same evidence everywhere."""
import pytest
import torch

from oracle import mopoe_oracle as mo, ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


@pytest.mark.parametrize("method", ["joint_elbo", "moe", "poe", "jsd"])
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


def test_rsa_oracle_equals_the_reference_stat_utils():
    """oracle/rsa_oracle.py against the unmodified experiments/stat_utils.py (data2cmat, vec2cmat, cmat2triu, fit_rsa)."""
    import numpy as np
    rh.install()
    import stat_utils as ref
    from oracle import rsa_oracle as ro
    rng = np.random.default_rng(5)
    lat = rng.standard_normal((40, 20)).astype(np.float32)
    score = np.round(rng.standard_normal(40) * 2).astype(np.float32)          # discrete: ties
    sex = rng.integers(0, 2, 40)
    assert np.array_equal(ro.data2cmat(lat), ref.data2cmat(lat))
    assert np.array_equal(ro.vec2cmat(score), ref.vec2cmat(score))
    assert np.array_equal(ro.vec2cmat(sex, categorical=True), ref.vec2cmat(sex, categorical=True))
    cm = ref.data2cmat(lat)
    assert np.array_equal(ro.cmat2triu(cm), ref.cmat2triu(cm))
    for other in (ref.vec2cmat(score), ref.vec2cmat(sex, categorical=True)):
        assert ro.fit_rsa(cm, other) == tuple(ref.fit_rsa(cm, other))
