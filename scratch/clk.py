import os, sys, time, torch, subprocess, threading, ctypes as C
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
from oracle import mopoe_oracle as mo
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, mo.init_params(mo.ModelSpec(**bench.HBN), seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
src, dst = src.cuda(), dst.cuda()
L = _lib.lib(); L.mopoe_profile_enable(1)
ws = engine.Workspace()
for mat in (True, False):
    r = None
    for i in range(3): r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, materialize=mat)
    torch.cuda.synchronize()
    s = bench.ClockSampler(0); s.start(); time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3000
    for i in range(n): r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, materialize=mat, out=r) if mat else daa.daa_sweep(spec, flat, src, dst, 150, 1000, seed=1037, workspace=ws, materialize=mat)
    e1.record(); torch.cuda.synchronize()
    ms = C.c_float(); L.mopoe_daa_last_kernel_ms(C.byref(ms))
    rows = list(s.rows)
    c = s.stop()
    print("materialize=%s: %.4f ms per sweep sustained over %d sweeps, last pipe kernel %.4f ms, clocks %s" % (mat, e0.elapsed_time(e1) / n, n, ms.value, c))
    print("   samples (sm MHz, W):", [(x[0], x[2]) for x in rows[3:40:3]])
