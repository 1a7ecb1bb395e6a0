"""BASELINE.json configs[2]: full HBN-shaped training, 550 epochs, lr 0.002, beta 1, fused fwd+bwd+Adam persistent
kernel on 1 x B200.  Records (a) the first epochs against the CPU oracle on the SAME batch plan and the SAME noise
(the production Philox draws materialised with the numpy restatement), (b) the whole run through the reference-shaped
entry point workflow.train_exp: wall time, loss curve, checkpoint round trip.  Output: one JSON document."""
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import mopoe_b200
from mopoe_b200 import data, engine, workflow, _lib
from oracle import mopoe_oracle as mo, philox, cases

out = {"config": "HBN-shaped synthetic cohort (2048 complete + 512 clinical-only + 256 rois-only train subjects, 512 test), joint_elbo, "
                 "input_dims [7,444], latent 20, style [3,20], batch 256, lr 0.002, beta 1, 550 epochs, missing blocks allowed"}
dev = torch.device("cuda")
# ---- (a) first epochs vs the oracle -----------------------------------------------------------------
spec = mopoe_b200.PathSpec(cases.HBN["dims"], cases.HBN["style_dims"], 20, "joint_elbo", cases.HBN["mod_names"])
ospec = mo.ModelSpec(**cases.HBN)
params0 = engine.init_params(spec, seed=0)
flat = engine.pack_params(spec, params0, dev)
cohort = data.make_cohort()
train = np.r_[0:2048, 2560:2560 + 512 + 256]
has = np.stack([cohort["has_clinical"][train], cohort["has_rois"][train]])
xs = [torch.from_numpy(cohort["clinical"][train]), torch.from_numpy(cohort["rois"][train])]
rng = np.random.RandomState(0)
plan = []
for _ in range(3):
    plan += data.epoch_plan(has, 256, rng)
n_steps, seed = len(plan), 4242
offs = np.cumsum([0] + [len(ix) for _, ix in plan])
index = torch.from_numpy(np.concatenate([ix for _, ix in plan]).astype(np.int32)).to(dev)
bdev = engine.make_batches(spec, [(len(ix), mask, int(offs[i])) for i, (mask, ix) in enumerate(plan)], dev)
m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
t_ = torch.zeros(4, dtype=torch.int32, device=dev)
sc = engine.train_steps(spec, flat, [x.to(dev) for x in xs], bdev, n_steps, 256, 2, row_index=[index, index], seed=seed,
                        adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002).cpu().numpy()
E = spec.eps_width
eps_all = torch.from_numpy(philox.philox_normal(seed, philox.STREAM_TRAIN, n_steps * 256 * E)).view(n_steps, 1, 256, E)
params = {k: v.clone() for k, v in params0.items()}
opt = mo.Adam(params, lr=0.002)
worst, olosses = 0.0, []
for i, (mask, ix) in enumerate(plan):
    ixl = torch.from_numpy(ix.astype(np.int64))
    batch = {n: xs[m][ixl] for m, n in enumerate(spec.mod_names) if mask >> m & 1}
    o, g, used = mo.elbo_and_grads(params, ospec, batch, eps_all[i][:, :len(ix)])
    params = opt.step(params, g, used)
    olosses.append(float(o["total_loss"]))
    worst = max(worst, abs(sc[i, 0] - olosses[-1]) / abs(olosses[-1]))
got = engine.unpack_params(spec, flat)
perr = max(float((got[k].cpu() - params[k]).abs().max() / params[k].abs().max()) for k in params)
out["first_epochs_vs_oracle"] = {"steps": n_steps, "noise": "in-kernel Philox == numpy restatement fed to the oracle",
                                 "max_rel_loss_diff": worst, "max_rel_param_diff_after": perr,
                                 "gpu_loss_first_last": [float(sc[0, 0]), float(sc[-1, 0])], "oracle_loss_first_last": [olosses[0], olosses[-1]]}
# ---- (b) the 550-epoch run through train_exp --------------------------------------------------------
tmp = tempfile.mkdtemp()
ds, outdir = os.path.join(tmp, "data"), os.path.join(tmp, "out")
os.makedirs(outdir)
data.write_dataset(ds, data.make_cohort(standardize=False))
torch.cuda.synchronize()
t0 = time.perf_counter()
run = workflow.train_exp("hbn", ds, outdir, [7, 444], num_epochs=550, batch_size=256, learning_rate=0.002, beta=1.0,
                         method="joint_elbo", data_seed=3)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
rundir = os.path.join(outdir, run)
tr = np.load(os.path.join(rundir, "logs", "scalars_train_model0.npy"))
te = np.load(os.path.join(rundir, "logs", "scalars_test_model0.npy"))
steps_per_epoch = len(tr) // 550
tr_epoch = tr[:, 0].reshape(550, steps_per_epoch)
full = tr[:, _lib.N_SCALARS - 17].reshape(550, steps_per_epoch) if False else None
te_epoch = te[:, 0].reshape(550, -1).mean(1)
rows = float(tr[:, 45].sum()) if tr.shape[1] > 45 else None
sd = torch.load(os.path.join(rundir, "checkpoints", "0549", "model"))
out["run_550_epochs"] = {"wall_s": wall, "train_steps": int(len(tr)), "steps_per_epoch": int(steps_per_epoch),
                         "impl": ["cuda-core", "tcgen05"][_lib.lib().mopoe_train_last_impl()],
                         "includes": "epoch plans (numpy), 550 train + 550 test launches, 110 checkpoints (torch.save), flags.rar",
                         "train_loss_epoch_mean": {str(e): float(tr_epoch[e].mean()) for e in (0, 1, 5, 10, 25, 50, 100, 200, 300, 400, 549)},
                         "test_loss_epoch_mean": {str(e): float(te_epoch[e]) for e in (0, 1, 5, 10, 25, 50, 100, 200, 300, 400, 549)},
                         "all_finite": bool(np.isfinite(tr).all() and np.isfinite(te).all()),
                         "checkpoint_keys_match_reference": list(sd.keys()) == list(mo.param_shapes(ospec).keys()),
                         "reference_cpu_estimate_s": "9.5 ms per step on 8 host cores (oracle port) x %d steps = %.0f s" % (len(tr), 9.5e-3 * len(tr))}
print(json.dumps(out, default=float))
