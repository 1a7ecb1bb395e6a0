import sys, numpy as np, torch, ctypes as C
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import engine, _lib
from oracle import cases, mopoe_oracle as mo
device = torch.device("cuda")
S = cases.STRESS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = 4
method = "joint_elbo"
spec = mopoe_b200.PathSpec(S["dims"], S["style_dims"], S["latent_dim"], method, S["mod_names"])
flat = engine.pack_params(spec, mo.init_params(mo.ModelSpec(**dict(S, method=method)), seed=0), device)
g = torch.Generator(device="cuda").manual_seed(0)
data = [torch.randn(N, d, device=device, generator=g) for d in S["dims"]]
bd = engine.make_batches(spec, [(N, 15, 0)] * steps, device)
m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
t_ = torch.zeros(4, dtype=torch.int32, device=device)
ws = engine.Workspace()
kw = dict(seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
L = _lib.lib(); buf = (C.c_float * 16)()
sc = engine.train_steps(spec, flat, data, bd, steps, N, 2, **kw); torch.cuda.synchronize()
L.mopoe_debug_p2prof(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sc = engine.train_steps(spec, flat, data, bd, steps, N, 2, **kw); e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / steps)
L.mopoe_debug_p2prof(buf)
ph = sc.cpu()[1:, 56:62].mean(0).tolist()
for n, c in zip(["P1 (cta0 work)", "barrier1 wait", "P2 (cta0 work)", "barrier2 wait", "P3 (cta0 work)", "barrier3 wait"], ph): print("%-18s %10.0f cycles %8.1f us" % (n, c, c / 1965.0))
names = ["hidden->smem", "heads", "latent fwd", "style fwd", "decoders+dx+dz", "latent bwd", "style bwd", "dA"]
for n, c in zip(names, list(buf)[:8]): print("  P2 %-16s %10.0f cycles/step %8.1f us (cta 0, all its tiles)" % (n, c / steps, c / steps / 1965.0))
