"""Per-stage cycles of the CUDA-core training tile (P2) and per-phase cycles of the step, from a -DTRAIN_PROF build:
  make -C 2022_cambroise_interpret_multivae_b200/csrc EXTRA=-DTRAIN_PROF OBJDIR=/tmp/tprof OUT=/root/repo/scratch/variants/lib_tprof.so
"""
import ctypes as C, sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import engine, _lib
_lib.LIB_PATH = "/root/repo/scratch/variants/lib_tprof.so"
import bench, numpy as np
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, sys.argv[1] if len(sys.argv) > 1 else "joint_elbo", bench.HBN["mod_names"])
dev = torch.device("cuda")
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), dev)
g = torch.Generator().manual_seed(0)
n, k = 256, 300
dd = [torch.randn(n, d, generator=g).to(dev) for d in spec.dims]
idx = torch.arange(n, dtype=torch.int32, device=dev)
bdev = engine.make_batches(spec, [(n, 3, 0)] * k, dev)
m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
t_ = torch.zeros(4, dtype=torch.int32, device=dev)
ws = engine.Workspace()
go = lambda: engine.train_steps(spec, flat, dd, bdev, k, n, 2, row_index=[idx] * 2, seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
go(); torch.cuda.synchronize()
L = _lib.lib()
buf = (C.c_float * 16)()
L.mopoe_debug_p2prof(buf)            # clear
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sc = go(); e1.record(); torch.cuda.synchronize()
L.mopoe_debug_p2prof(buf)
p2 = np.array(list(buf)) / k
names = ["stage h rows", "heads", "lat_forward", "sync", "decoders + nll + dz", "lat_backward", "sync", "d heads / dA"]
print("us per step %.2f" % (1e3 * e0.elapsed_time(e1) / k))
for i, nm in enumerate(names):
    print("P2 %-22s %7.0f cycles / step (CTA 0)" % (nm, p2[i]))
ph = sc.cpu().numpy()[:, 56:62].mean(0)
for nm, v in zip(["P1 work", "P1 barrier", "P2 work", "P2 barrier", "P3 work", "P3 barrier"], ph):
    print("%-12s %7.0f cycles" % (nm, v))
