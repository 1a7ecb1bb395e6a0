"""Per-role cycle counters of a -DPK_PROF build of the avatar pipe kernel (scratch/variants/lib_prof.so), mean over CTAs,
per tile.  Slots: [0..8) producer warp 0, [8..16) producer warp 4, [16..20) epilogue warp 0 (work, wait acc_full, total),
[20..24) decoder issuer (issue, issue, waits), [24..32) aux (publish/fold, wait z_full, prologue, CTA cycles, timers)."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import daa, engine, _lib
_lib.LIB_PATH = "/root/repo/scratch/variants/lib_%s.so" % (sys.argv[1] if len(sys.argv) > 1 else "prof")
import bench, numpy as np
spec = mopoe_b200.PathSpec(bench.HBN["dims"], bench.HBN["style_dims"], 20, "joint_elbo", bench.HBN["mod_names"])
flat = engine.pack_params(spec, engine.init_params(spec, seed=0), torch.device("cuda"))
src, dst = bench.draw_validation_batches(20, 1037)
ws = engine.Workspace()
for i in range(4):
    r = daa.daa_sweep(spec, flat, src.cuda(), dst.cuda(), 150, 1000, workspace=ws, base_mean="direct")
rows = []
for cta in range(148):
    os.environ["MOPOE_PHASE_CTA"] = str(cta)
    rows.append(daa.phase_cycles(spec, r))
a = np.array(rows, dtype=np.float64)
tiles = 20 * 411 / 148.0
m = a.mean(0) / tiles
names = {0: "prod0 wait cache(aux)", 1: "prod0 P1", 5: "prod0 wait z_free", 2: "prod0 style chunks + noise", 3: "prod0 wait heads_done",
         4: "prod0 content chunks", 6: "prod0 arrive", 8: "prod4 wait cache", 9: "prod4 P1", 13: "prod4 wait z_free", 10: "prod4 style+noise",
         11: "prod4 wait heads_done", 12: "prod4 content chunks", 16: "epi work", 17: "epi wait acc_full", 20: "dmma after z_full",
         21: "dmma issue", 22: "dmma waits (z_full + acc_empty)", 24: "aux publish", 25: "aux wait z_full"}
for k in sorted(names):
    print("%-34s %8.0f cycles / tile" % (names[k], m[k]))
print("CTA cycles / tile %.0f (mean over CTAs), min %.0f max %.0f" % (a[:, 27].mean() / tiles, a[:, 27].min() / tiles, a[:, 27].max() / tiles))
