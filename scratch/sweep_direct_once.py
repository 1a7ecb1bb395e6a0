"""A few direct-mean HBN sweeps (for ncu captures of the small kernels of the sweep)."""
import sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
src, dst = src.cuda(), dst.cuda()
ws = engine.Workspace()
r = None
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, out=r, base_mean="direct")
torch.cuda.synchronize()
print("ok")
