"""Per-stage cycle counters of the tensor-core training kernel (profiling build, MOPOE_LIB_PATH=...prof.so)."""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
os.environ["MOPOE_LIB_PATH"] = "/root/repo/2022_cambroise_interpret_multivae_b200/libmopoe_b200_prof.so"
os.environ["MOPOE_TRAIN_IMPL"] = "tc"
import mopoe_b200
from mopoe_b200 import engine, _lib
from oracle import cases
NAMES = {0: "setup", 1: "x convert (+wait xempty)", 2: "wait P1 acc", 3: "P1 epilogue", 4: "wait S1 acc", 5: "S1 epilogue", 6: "latent fwd",
         7: "zop convert", 8: "wait S2 acc", 9: "S2 epilogue", 10: "wait dz acc", 11: "dz epilogue", 12: "latent bwd", 13: "deop convert",
         14: "wait S4 acc", 15: "S4 epilogue", 30: "loader: ring full (wait empty)", 31: "mma: wait chunk (ring_full)", 32: "mma: wait x block", 33: "mma: wait B operand", 40: "lat: loads+exp", 41: "lat: owner", 42: "lat: subsets", 43: "lat: noise+z", 44: "lat: class reductions", 45: "lat: style loop", 46: "lat: style reductions", 34: "P3 loader: ring full", 35: "P3 mma: wait chunk", 36: "P3 compute: wait acc", 37: "P3 epilogue", 20: "prep", 21: "barrier 1", 22: "P2 (CTA 0)", 23: "barrier 2", 24: "P3 (CTA 0)", 25: "barrier 3"}
def prof(base, method, n, steps, inject=False):
    spec = mopoe_b200.PathSpec(base["dims"], base["style_dims"], base["latent_dim"], method, base["mod_names"])
    dev = torch.device("cuda")
    flat = engine.pack_params(spec, engine.init_params(spec, seed=0), dev)
    g = torch.Generator().manual_seed(0)
    rows = max(n, 4096)
    dd = [torch.randn(rows, d, generator=g).to(dev) for d in spec.dims]
    idx = torch.from_numpy(np.concatenate([np.random.RandomState(s).permutation(rows)[:n] for s in range(steps)]).astype(np.int32)).to(dev)
    bdev = engine.make_batches(spec, [(n, (1 << spec.n_mods) - 1, s * n) for s in range(steps)], dev)
    m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
    t_ = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = engine.Workspace()
    eps = torch.randn(steps, spec.n_pass, n, spec.eps_width, device=dev) if inject else None
    go = lambda: engine.train_steps(spec, flat, dd, bdev, steps, n, 2, row_index=[idx] * spec.n_mods, seed=7, eps=eps, adam_m=m_, adam_v=v_, adam_t=t_, workspace=ws)
    go(); torch.cuda.synchronize()
    out = (C.c_float * 64)()
    L = _lib.lib()
    L.mopoe_debug_tcprof.argtypes = [C.c_void_p]
    L.mopoe_debug_tcprof(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    L.mopoe_debug_tcprof(out)
    print("== inject=%s %s %s n=%d: %.1f us/step" % (inject, "hbn" if base is cases.HBN else "stress", method, n, 1e3 * e0.elapsed_time(e1) / steps))
    for i, nm in NAMES.items():
        print("   %-28s %8.2f us/step" % (nm, out[i] / steps / 1965.0))
prof(cases.HBN, "joint_elbo", 256, 200)
prof(cases.STRESS, "joint_elbo", 65536, 3)
