"""A/B of environment switches on the direct-mean sweep: python scratch/ab_direct.py VAR=1 [VAR2=1 ...] (each in its own process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import mopoe_b200
from mopoe_b200 import daa, engine
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
src, dst = src.cuda(), dst.cuda()
ws = engine.Workspace()
r = None
for i in range(5): r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, out=r, base_mean="direct")
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, out=r, base_mean="direct")
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(g): r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, out=r, base_mean="direct")
for i in range(5): g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
import hashlib
print("%%-40s sweep (graph) %%.4f ms  coefs md5 %%s" %% (sys.argv[1], e0.elapsed_time(e1) / 50, hashlib.md5(r.coefs.cpu().numpy().tobytes()).hexdigest()[:8]))
''' % ROOT
for setting in [""] + sys.argv[1:]:
    env = dict(os.environ)
    for kv in setting.split(","):
        if kv:
            k, v = kv.split("=")
            env[k] = v
    subprocess.run([sys.executable, "-c", CHILD, setting or "default"], env=env, check=False)
