import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200, bench
from mopoe_b200 import data, engine

device = torch.device("cuda")
spec_kw = bench.HBN
cohort = data.make_cohort()
train = np.r_[0:2048]
dev_data = [torch.from_numpy(cohort["clinical"][train]).to(device), torch.from_numpy(cohort["rois"][train]).to(device)]
steps = 300
for method in ("joint_elbo",):
    spec = mopoe_b200.PathSpec(spec_kw["dims"], spec_kw["style_dims"], spec_kw["latent_dim"], method, spec_kw["mod_names"])
    flat = engine.pack_params(spec, engine.init_params(spec, seed=0), device)
    for N in (256,):
        nb = 2048 // N
        plan = [(N, 3, (i % nb) * N) for i in range(steps)]
        bd = engine.make_batches(spec, plan, device)
        index = torch.arange(2048, dtype=torch.int32, device=device)
        m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
        t_ = torch.zeros(4, dtype=torch.int32, device=device)
        g_ = torch.zeros_like(flat)
        ws = engine.Workspace()
        for mode in (0, 2):
            f = flat.clone()
            kw = dict(row_index=[index, index], seed=7, workspace=ws)
            if mode == 1: kw["grads"] = g_
            if mode == 2: kw.update(adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002)
            engine.train_steps(spec, f, dev_data, bd, steps, N, mode, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.train_steps(spec, f, dev_data, bd, steps, N, mode, **kw)
            e1.record(); torch.cuda.synchronize()
            print(method, "N", N, "mode", mode, "us/step %.1f" % (1e3 * e0.elapsed_time(e1) / steps))
sc = engine.train_steps(spec, f, dev_data, bd, steps, N, 2, **kw).cpu()
ph = sc[50:, 56:62].mean(0).tolist()
names = ["P1 (cta0 work)", "barrier1 wait", "P2 (cta0 work)", "barrier2 wait", "P3 (cta0 work)", "barrier3 wait"]
for n, c in zip(names, ph): print("%-18s %8.0f cycles %6.1f us" % (n, c, c / 1965.0))
import ctypes as C
from mopoe_b200 import _lib
L = _lib.lib()
buf = (C.c_float * 16)()
L.mopoe_debug_p2prof(buf)   # reset
sc = engine.train_steps(spec, f, dev_data, bd, steps, N, 2, **kw); torch.cuda.synchronize()
L.mopoe_debug_p2prof(buf)
names = ["hidden->smem", "heads", "latent fwd", "style fwd", "decoders+dx+dz", "latent bwd", "style bwd", "dA"]
for n, c in zip(names, list(buf)[:8]): print("  P2 %-16s %8.0f cycles/step %6.1f us" % (n, c / steps, c / steps / 1965.0))
