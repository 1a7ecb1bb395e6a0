"""A/B timing of library variants: python scratch/ab.py [variant names...]  ('' = product library).
Each variant runs in its own process (ctypes cannot reload)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, ctypes as C, torch
sys.path.insert(0, %r)
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
if sys.argv[1]: _lib.LIB_PATH = sys.argv[1]

import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
src, dst = src.cuda(), dst.cuda()
L = _lib.lib(); L.mopoe_profile_enable(1)
ws = engine.Workspace()
r = None
ts = []
for i in range(8):
    r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, out=r)
    ms = C.c_float(); torch.cuda.synchronize(); L.mopoe_daa_last_kernel_ms(C.byref(ms)); ts.append(ms.value)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(10): r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, out=r)
e1.record(); torch.cuda.synchronize()
import hashlib
h = hashlib.md5(r.coefs.cpu().numpy().tobytes()).hexdigest()[:8]
print("%%-24s pipe kernel %%.4f ms   sweep %%.4f ms   coefs md5 %%s nan=%%d" %% (sys.argv[2], sorted(ts)[len(ts)//2], e0.elapsed_time(e1)/10, h, int(torch.isnan(r.pvalues).sum())))
''' % ROOT
for name in (sys.argv[1:] or [""]):
    path = os.path.join(ROOT, "scratch", "variants", "lib_%s.so" % name) if name else ""
    subprocess.run([sys.executable, "-c", CHILD, path, name or "product"], check=False)
