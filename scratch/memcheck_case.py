import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
from oracle import cases, mopoe_oracle as mo
case = dict(cases._case(cases.HBN, "joint_elbo", True, (0, 1), 5, 73, 173), n_val=2, n_base=7, n_samples=131, sample_latents=True)
ospec = cases.spec_of(case)
spec = mopoe_b200.PathSpec(ospec.dims, ospec.style_dims, ospec.latent_dim, ospec.method, ospec.mod_names)
flat = engine.pack_params(spec, mo.init_params(ospec, seed=1), torch.device("cuda"))
src, dst, eb, es, ea = cases.daa_inputs_of(case, ospec)
r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 131, 7, seed=3)
torch.cuda.synchronize()
assert _lib.lib().mopoe_daa_last_impl() == 2
print("pipe ok", float(r.coefs.abs().sum()), bool(torch.isfinite(r.avatars).all()))
# one fused training launch (2 steps, small batch -> 1-row tiles, split-K P1, 2 CTAs per SM)
data = [torch.randn(64, d, device="cuda") for d in ospec.dims]
bd = engine.make_batches(spec, [(33, 3, 0), (31, 3, 33)], torch.device("cuda"))
m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
t_ = torch.zeros(4, dtype=torch.int32, device="cuda")
sc = engine.train_steps(spec, flat.clone(), data, bd, 2, 64, 2, seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002)
torch.cuda.synchronize()
print("train ok", sc[:, 0].tolist())
