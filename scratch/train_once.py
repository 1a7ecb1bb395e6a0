"""One launch of the small-batch training kernel (256 rows, 100 steps) -- the target of an ncu capture."""
import sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import engine
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
dev = torch.device("cuda")
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), dev)
g = torch.Generator().manual_seed(0)
n, k = 256, 100
dd = [torch.randn(n, d, generator=g).to(dev) for d in spec.dims]
idx = torch.arange(n, dtype=torch.int32, device=dev)
bdev = engine.make_batches(spec, [(n, 3, 0)] * k, dev)
m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
t_ = torch.zeros(4, dtype=torch.int32, device=dev)
ws = engine.Workspace()
for _ in range(2):
    engine.train_steps(spec, flat, dd, bdev, k, n, 2, row_index=[idx] * 2, seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
torch.cuda.synchronize()
print("ok")
