import os, sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
_lib.LIB_PATH = "/root/repo/scratch/variants/lib_%s.so" % (sys.argv[1] if len(sys.argv) > 1 else "prof")
from oracle import mopoe_oracle as mo
import bench
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, mo.init_params(mo.ModelSpec(**bench.HBN), seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
ws = engine.Workspace()
for i in range(3):
    r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 150, 1000, workspace=ws, materialize=os.environ.get("MAT", "1") == "1")
for cta in ("0", "73", "147"):
    os.environ["MOPOE_PHASE_CTA"] = cta
    ph = daa.phase_cycles(spec, r)
    names = ["w0 cache wait", "w0 P1", "w0 noise->heads wait", "w0 heads wait", "w0 passes rest", "w0 z_free wait", "w0 arrive", "-",
             "w4 cache wait", "w4 P1", "w4 noise->heads wait", "w4 heads wait", "w4 passes rest", "w4 z_free wait", "w4 arrive", "-",
             "E wait acc_full (pc0)", "E drain+store work (pc1)", "E total", "-", "D pc0 (after z_full wait)", "D pc1 (issue after acc_empty)", "D pc2 (waits)", "-", "AUX work", "AUX wait"]
    print("CTA", cta)
    for n, c in zip(names, ph):
        if n != "-": print("  %-24s %9d cycles  per tile %7.0f" % (n, c, c / 56))
import ctypes as C
L = _lib.lib(); L.mopoe_profile_enable(1)
ts = []
for i in range(5):
    r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 150, 1000, workspace=ws, materialize=os.environ.get("MAT", "1") == "1")
    ms = C.c_float(); torch.cuda.synchronize(); L.mopoe_daa_last_kernel_ms(C.byref(ms)); ts.append(ms.value)
print("kernel ms", sorted(ts))
