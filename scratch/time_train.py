"""us per fused fwd+bwd+Adam step: HBN batch 256 (300 steps per launch) and the stress shape (batch 65 536)."""
import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import engine, data, _lib
from oracle import cases

def timeit(base, method, n, steps, impl):
    os.environ["MOPOE_TRAIN_IMPL"] = impl
    spec = mopoe_b200.PathSpec(base["dims"], base["style_dims"], base["latent_dim"], method, base["mod_names"])
    dev = torch.device("cuda")
    flat = engine.pack_params(spec, engine.init_params(spec, seed=0), dev)
    g = torch.Generator().manual_seed(0)
    rows = max(n, 4096)
    dd = [torch.randn(rows, d, generator=g).to(dev) for d in spec.dims]
    idx = torch.from_numpy(np.concatenate([np.random.RandomState(s).permutation(rows)[:n] for s in range(steps)]).astype(np.int32)).to(dev)
    full = (1 << spec.n_mods) - 1
    bdev = engine.make_batches(spec, [(n, full, s * n) for s in range(steps)], dev)
    m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
    t_ = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = engine.Workspace()
    def go():
        return engine.train_steps(spec, flat, dd, bdev, steps, n, 2, row_index=[idx] * spec.n_mods, seed=7, adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
    go(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc = go(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-6s %-10s n=%-6d impl=%s(%d)  %.1f us/step  (%.3g samples/s) loss %.3f" % (
        "hbn" if base is cases.HBN else "stress", method, n, impl, _lib.lib().mopoe_train_last_impl(), 1e3 * ms / steps, n * steps / (ms * 1e-3), float(sc[-1, 0])), flush=True)

for impl in sys.argv[1:] or ["tc", "ffma"]:
    for method in ("joint_elbo", "moe", "poe"):
        timeit(cases.HBN, method, 256, 300, impl)
    timeit(cases.HBN, "joint_elbo", 4096, 20, impl)
    timeit(cases.HBN, "joint_elbo", 65536, 5, impl)
    timeit(cases.STRESS, "joint_elbo", 65536, 3, impl)
    timeit(cases.STRESS, "poe", 65536, 3, impl)
