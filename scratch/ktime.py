import sys, ctypes as C, torch
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
def run(**kw):
    ts = []
    for i in range(6):
        r = daa.daa_sweep(spec, flat, src, dst, 150, 1000, workspace=ws, **kw)
        ms = C.c_float(); torch.cuda.synchronize(); L.mopoe_daa_last_kernel_ms(C.byref(ms)); ts.append(ms.value)
    return sorted(ts)[len(ts)//2]
print("materialize=True  kernel ms", run())
print("materialize=False kernel ms", run(materialize=False))
