import os, sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
_lib.LIB_PATH = "/root/repo/scratch/variants/lib_prof.so"

import bench, numpy as np, ctypes as C
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
ws = engine.Workspace()
L = _lib.lib(); L.mopoe_profile_enable(1)
for i in range(4):
    r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 150, 1000, workspace=ws)
ms = C.c_float(); torch.cuda.synchronize(); L.mopoe_daa_last_kernel_ms(C.byref(ms))
rows = []
for cta in range(148):
    os.environ["MOPOE_PHASE_CTA"] = str(cta)
    ph = daa.phase_cycles(spec, r)
    rows.append(ph[24:32])
a = np.array(rows, dtype=np.float64)
print("event kernel ms %.4f" % ms.value)
print("prologue cycles: min %.0f mean %.0f max %.0f" % (a[:, 2].min(), a[:, 2].mean(), a[:, 2].max()))
print("CTA cycles:      min %.0f mean %.0f max %.0f" % (a[:, 3].min(), a[:, 3].mean(), a[:, 3].max()))
g0, g1 = a[:, 4], a[:, 5]
print("globaltimer: first entry -> last entry %.1f us, first entry -> last exit %.1f us, first exit -> last exit %.1f us" % ((g0.max() - g0.min()) / 1e3, (g1.max() - g0.min()) / 1e3, (g1.max() - g1.min()) / 1e3))
