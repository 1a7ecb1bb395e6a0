import os, sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine
from oracle import mopoe_oracle as mo
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, mo.init_params(mo.ModelSpec(**bench.HBN), seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
ws = engine.Workspace()
for i in range(3):
    r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 150, 1000, workspace=ws, materialize=os.environ.get("MAT", "1") == "1")
ph = daa.phase_cycles(spec, r)
import os
print("CTA", os.environ.get("MOPOE_PHASE_CTA", "max"))
names = ["P X-barrier(issue)", "P butterfly (inside passes)", "P wait heads(i)", "P p1(i+1)", "P passes (posterior..)", "P wait e_full", "P arrive", "P tmem heads ld+wait (inside passes)",
         "P.w4 0", "P.w4 1", "P.w4 2", "P.w4 3", "P.w4 4", "P.w4 5", "P.w4 6", "-",
         "E wait acc_full", "E drain+store", "E total", "-", "D wait z_full", "D wait acc_empty", "D issue", "-",
         "AUX X-barrier(issue)", "AUX work+wait"]
for n, c in zip(names, ph): print("%-28s %9d cycles  per tile %7.0f" % (n, c, c/56))
