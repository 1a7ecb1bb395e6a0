import os, sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
_lib.LIB_PATH = "/root/repo/scratch/variants/lib_prof.so"
from oracle import mopoe_oracle as mo
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, mo.init_params(mo.ModelSpec(**bench.HBN), seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
ws = engine.Workspace()
for i in range(3):
    r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 150, 1000, workspace=ws)
tot = []
for cta in range(148):
    os.environ["MOPOE_PHASE_CTA"] = str(cta)
    ph = daa.phase_cycles(spec, r)
    tot.append(ph[18])
import numpy as np
t = np.array(tot) / 1e3
print("E total per CTA (K cycles): min %.0f max %.0f mean %.0f" % (t.min(), t.max(), t.mean()))
print(" ".join("%d" % x for x in t))
