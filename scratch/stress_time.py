import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import engine
from oracle import cases, mopoe_oracle as mo
device = torch.device("cuda")
S = cases.STRESS
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = 10
for method in ("joint_elbo", "poe", "moe"):
    spec = mopoe_b200.PathSpec(S["dims"], S["style_dims"], S["latent_dim"], method, S["mod_names"])
    flat = engine.pack_params(spec, mo.init_params(mo.ModelSpec(**dict(S, method=method)), seed=0), device)
    g = torch.Generator(device="cuda").manual_seed(0)
    data = [torch.randn(N, d, device=device, generator=g) for d in S["dims"]]
    bd = engine.make_batches(spec, [(N, 15, 0)] * steps, device)
    m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
    t_ = torch.zeros(4, dtype=torch.int32, device=device)
    ws = engine.Workspace()
    kw = dict(seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
    sc = engine.train_steps(spec, flat, data, bd, steps, N, 2, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc = engine.train_steps(spec, flat, data, bd, steps, N, 2, **kw); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    macs = 0
    for d, s in zip(S["dims"], S["style_dims"]):
        macs += d * 256 + 256 * (40 + 2 * s) + (20 + s) * d
    flop = 2 * macs * 3 * N * (1 + (len(S["dims"]) if method == "poe" else 0) * 0)   # fwd + dW + dX ~ 3x fwd
    print(method, "N", N, "ms/step %.3f" % ms, "samples/s %.3e" % (N / ms * 1e3), "approx TFLOP/s %.1f" % (flop / ms / 1e9), "loss", float(sc[-1, 0]))
