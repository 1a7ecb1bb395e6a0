"""First light of the tensor-core training kernel: loss terms and per-parameter gradient errors vs the oracle."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
os.environ["MOPOE_TRAIN_IMPL"] = sys.argv[1] if len(sys.argv) > 1 else "tc"
import mopoe_b200
from mopoe_b200 import engine, _lib
from oracle import cases, mopoe_oracle as mo

def run(base, method, n, present=None, fact=True, mode=1):
    present = present or tuple(range(len(base["dims"])))
    case = cases._case(base, method, fact, present, n, 300 + n % 97, 400 + n % 89)
    ospec = cases.spec_of(case)
    spec = mopoe_b200.PathSpec(ospec.dims, ospec.style_dims, ospec.latent_dim, ospec.method, ospec.mod_names)
    params = mo.init_params(ospec, seed=case["seed"])
    flat = engine.pack_params(spec, params, torch.device("cuda"))
    batch, eps = cases.inputs_of(case, ospec)
    grads = torch.zeros_like(flat)
    data = [batch[k].cuda().contiguous() if k in batch else None for k in spec.mod_names]
    bdev = engine.make_batches(spec, [(n, spec.present_mask(batch.keys()), 0)], flat.device)
    t0 = time.time()
    sc = engine.train_steps(spec, flat, data, bdev, 1, n, mode, eps=eps.cuda().contiguous()[None], grads=grads)
    torch.cuda.synchronize()
    dt = time.time() - t0
    out, g, used = mo.elbo_and_grads(params, ospec, batch, eps)
    got = engine.unpack_params(spec, grads)
    tag = "%s %s n=%d present=%s impl=%d" % ("hbn" if base is cases.HBN else "stress", method, n, present, _lib.lib().mopoe_train_last_impl())
    print(tag, "loss got %.6f want %.6f  (%.2fs)" % (float(sc[0, 0]), float(out["total_loss"]), dt), flush=True)
    sc = sc.cpu().numpy()
    for k, v in out["log_probs"].items():
        print("   nll", k, sc[0, _lib.S_NLL + spec.mod_names.index(k)], float(v))
    print("   joint_div", sc[0, _lib.S_JOINT_DIV], float(out["joint_divergence"]))
    if mode == 0: return
    for k in g:
        if not used[k]: continue
        a, w = got[k].cpu().double(), g[k].double()
        rel = float((a - w).abs().max() / w.abs().max())
        flag = "" if rel < 1e-4 else "   <<<<<"
        print("   %-45s rel %.2e  |want| %.3e%s" % (k, rel, float(w.abs().max()), flag))

torch.cuda.init()
run(cases.HBN, "joint_elbo", 16)
run(cases.HBN, "joint_elbo", 256)
run(cases.HBN, "joint_elbo", 37)
run(cases.HBN, "poe", 96)
run(cases.HBN, "moe", 96, present=(1,))
run(cases.STRESS, "joint_elbo", 96)
run(cases.HBN, "joint_elbo", 4097)
run(cases.STRESS, "poe", 4097)
