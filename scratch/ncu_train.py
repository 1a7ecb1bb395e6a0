"""One launch of the tensor-core training kernel on the stress shape (batch 65 536, 4 modalities, 15 subsets,
2 fused steps) -- the subject of the ncu --set full capture under profiles/."""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
os.environ["MOPOE_TRAIN_IMPL"] = "tc"
import mopoe_b200
from mopoe_b200 import engine, _lib
S = dict(dims=[7, 444, 24, 148], style_dims=[3, 20, 3, 20], latent_dim=20, mod_names=["clinical", "rois", "modc", "modd"])
spec = mopoe_b200.PathSpec(S["dims"], S["style_dims"], 20, "joint_elbo", S["mod_names"])
dev = torch.device("cuda")
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), dev)
g = torch.Generator().manual_seed(0)
n, k = 65536, 2
dd = [torch.randn(n, d, generator=g).to(dev) for d in spec.dims]
idx = torch.arange(n, dtype=torch.int32, device=dev)
bdev = engine.make_batches(spec, [(n, 15, 0)] * k, dev)
m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
t_ = torch.zeros(4, dtype=torch.int32, device=dev)
sc = engine.train_steps(spec, flat, dd, bdev, k, n, 2, row_index=[idx] * 4, seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002)
torch.cuda.synchronize()
assert _lib.lib().mopoe_train_last_impl() == 1 and bool(torch.isfinite(sc[:, 0]).all())
print("loss", sc[:, 0].tolist())
