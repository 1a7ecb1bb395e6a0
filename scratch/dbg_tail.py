"""Which gradient elements differ at large batches (tail tiles of the R=16 row tiling)?"""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import engine
from oracle import cases, mopoe_oracle as mo

def run(base, method, n, present=None):
    present = present or tuple(range(len(base["dims"])))
    case = cases._case(base, method, True, present, n, 300 + n % 97, 400 + n % 89)
    ospec = cases.spec_of(case)
    spec = mopoe_b200.PathSpec(ospec.dims, ospec.style_dims, ospec.latent_dim, ospec.method, ospec.mod_names)
    params = mo.init_params(ospec, seed=case["seed"])
    flat = engine.pack_params(spec, params, torch.device("cuda"))
    batch, eps = cases.inputs_of(case, ospec)
    grads = torch.zeros_like(flat)
    data = [batch[k].cuda().contiguous() if k in batch else None for k in spec.mod_names]
    bdev = engine.make_batches(spec, [(n, spec.present_mask(batch.keys()), 0)], flat.device)
    sc = engine.train_steps(spec, flat, data, bdev, 1, n, 1, eps=eps.cuda().contiguous()[None], grads=grads)
    torch.cuda.synchronize()
    out, g, used = mo.elbo_and_grads(params, ospec, batch, eps)
    got = engine.unpack_params(spec, grads)
    worst = []
    for k in g:
        if not used[k]: continue
        a, w = got[k].cpu().double(), g[k].double()
        err = (a - w).abs()
        rel = float(err.max() / w.abs().max())
        if rel > 5e-5:
            idx = np.unravel_index(int(err.argmax()), err.shape)
            nbad = int((err > 5e-5 * w.abs().max()).sum())
            worst.append((k, rel, idx, nbad, err.numel()))
    print(base is cases.HBN and "hbn" or "stress", method, n, "loss rel err %.2e" % abs(float(sc[0,0]) / float(out["total_loss"]) - 1))
    for w in worst: print("   ", w)

for n in (4096, 4097, 4100, 4112, 2049, 2064, 3000):
    run(cases.HBN, "joint_elbo", n)
run(cases.HBN, "moe", 65536)
run(cases.STRESS, "moe", 512)
run(cases.STRESS, "moe", 96)
run(cases.STRESS, "joint_elbo", 65536)
