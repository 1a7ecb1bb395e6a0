#!/usr/bin/env python
"""bench.py -- headline benchmark of the MoPoE-VAE hot path on B200.

metric   daa_avatars_per_s: digital avatars (one perturbed subject -> one 444-ROI vector) produced
         AND reduced to association statistics per second, whole job over all GPUs.
workload BASELINE.json configs[3]: HBN-shaped DAA sweep, n_validation=20 (per GPU: validations are
         the sharding unit, weak scaling), n_subjects=50, n_samples=150, M=1000 base passes,
         7 scores x 444 ROIs, joint_elbo, factorised default, hierarchical regression.
step     one full sweep of this rank's 20 validations: encoder pass, M base passes, score
         sampling, 52 500 avatar forwards per validation (warp-specialised tcgen05 pipeline), per-subject
         slopes, t-tests, and (N>1)
         the NCCL all_gather of the (n_val, 7, 444) fp64 coefs / p-value tables.
value    inputs resident in HBM, avatar tensor materialised in HBM (1.865 GB per sweep > L2).
e2e      same sweep through the public API with pinned HOST buffers: H2D of the drawn test batches,
         D2H of every array the reference's daa_exp writes (avatars included).
The `train` object (N=1 only) reports the fused fwd+bwd+Adam step (BASELINE.json configs[1-2] and the training
half of configs[4]) with its own roofline, e2e and CPU baseline (oracle port of run_epochs.train timed in the same
run); `--impl reference` carries the CPU training rate too.

`--impl reference` times the CPU port of the reference path (oracle/, torch CPU, all host threads)
on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBN = dict(dims=[7, 444], style_dims=[3, 20], latent_dim=20, mod_names=["clinical", "rois"])
DAA = dict(n_validation=20, n_subjects=50, n_samples=150, n_base=1000, seed=1037)
METRIC, UNIT = "daa_avatars_per_s", "avatars/s"
LAUNCH_NOTE = ["direct launches", "none"]


def workload_config(n_gpus):
    return {"workload": "HBN-shaped DAA sweep (BASELINE.json configs[3]): joint_elbo, input_dims [7,444], latent 20, "
                        "style [3,20], n_validation=%d per GPU, n_subjects=50, n_samples=150, M=1000, hierarchical "
                        "regression, 7 scores x 444 ROIs" % DAA["n_validation"],
            "n_validation_total": DAA["n_validation"] * n_gpus, "parallelism": "validations sharded over %d GPU(s)" % n_gpus,
            "l2": "outputs larger than L2: 1.865 GB avatar tensor written per sweep (126 MB L2)", "launch": LAUNCH_NOTE[0], "host_binding": LAUNCH_NOTE[1],
            "noise": "in-kernel philox (production mode); mean noise row of the M = 1000 affine base passes drawn directly as "
                     "N(0, 1/M) (base_mean='direct': same distribution, 1/M of the draws; the draw-for-draw M-pass mode is timed "
                     "beside it as value_base_draws)", "weights": "random init (seed 0)"}


def draw_validation_batches(n_val, seed, offset=0):
    """n_val batches of 50 test subjects with both blocks (workflow.py:362-372), host RNG."""
    from mopoe_b200 import data
    cohort = data.make_cohort()
    test = np.arange(2048, 2560)                      # the 512 complete test subjects
    rng = np.random.default_rng(seed + offset)
    idx = np.stack([rng.permutation(test)[:DAA["n_subjects"]] for _ in range(n_val)])
    return (torch.from_numpy(cohort["clinical"][idx]).contiguous(), torch.from_numpy(cohort["rois"][idx]).contiguous())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_daa_sample(budget_s):
    """Times the reference's DAA loop structure (one torch forward of batch 50 per base pass and per
    (sample, score); workflow.py:388-419) through the oracle port, then the closed-form hierarchical
    regression (the shipped reference uses statsmodels formula fits here: far slower).
    Returns (avatars, seconds, description)."""
    from oracle import daa_oracle, mopoe_oracle as mo
    torch.set_num_threads(os.cpu_count())
    spec = mo.ModelSpec(**HBN)
    params = mo.init_params(spec, seed=0)
    src, dst = draw_validation_batches(1, DAA["seed"])
    N, C_, E = DAA["n_subjects"], 7, spec.eps_width
    g = torch.Generator().manual_seed(0)
    # calibrate: cost of one forward
    with torch.no_grad():
        x = {"clinical": src[0], "rois": dst[0]}
        for _ in range(5):
            mo.forward(params, spec, x, torch.randn(N, E, generator=g))
        t0 = time.perf_counter()
        for _ in range(20):
            mo.forward(params, spec, x, torch.randn(N, E, generator=g))
        per_fwd = (time.perf_counter() - t0) / 20
    ratio = DAA["n_base"] / (DAA["n_samples"] * C_)
    n_samples = int(max(3, min(DAA["n_samples"], budget_s / per_fwd / (C_ * (1 + ratio)))))
    n_base = max(1, int(round(n_samples * C_ * ratio)))
    eb = torch.randn(1, n_base, N, E, generator=g)
    es = torch.randn(1, n_samples, N, C_, generator=g)
    ea = torch.randn(1, n_samples, C_, N, E, generator=g)
    t0 = time.perf_counter()
    av, sc, rc = daa_oracle.daa_generate(params, spec, src, dst, eb, es, ea)
    daa_oracle.hierarchical_regression(av, sc)
    dt = time.perf_counter() - t0
    desc = ("1 validation of 50 subjects, %d base passes + %d samples x 7 scores = %d torch-CPU forwards of batch 50 "
            "+ closed-form hierarchical regression (7x444 series)" % (n_base, n_samples, n_base + n_samples * C_))
    return n_samples * C_ * N, dt, desc


def run_reference(args, rank):
    if rank != 0:
        return
    budget = min(8.0, 150.0 / max(1, args.steps + args.warmup))
    times, avatars, desc = [], 0, ""
    for i in range(args.warmup + args.steps):
        n, dt, desc = cpu_daa_sample(budget)
        if i >= args.warmup:
            times.append(dt); avatars += n
    total = sum(times)
    value = avatars / total
    train = {}
    for method in ("joint_elbo",):
        rows, dt, st, tdesc = cpu_train_sample(min(10.0, 60.0 / max(1, args.steps)), method)
        train[method] = {"samples_per_s": rows / dt, "us_per_step": 1e6 * dt / max(1, st), "cores": os.cpu_count(),
                         "kind": "port", "sample": tdesc}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(times)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "statistics_parity": "unpinned: statsmodels (stat_utils.make_regression) is absent; closed forms checked against scipy/numpy",
            "train": train}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
STRESS = dict(dims=[7, 444, 24, 148], style_dims=[3, 20, 3, 20], latent_dim=20, mod_names=["clinical", "rois", "modc", "modd"])


def model_flops_per_row(spec_kw, method):
    """Algorithmic FLOPs of one training row: forward MACs of every modality (first layer, heads, decoder),
    x 2 FLOP, x 3 for forward + both backward contractions; poe decodes every modality twice."""
    L = spec_kw["latent_dim"]
    macs = 0
    for D, S in zip(spec_kw["dims"], spec_kw["style_dims"]):
        dec = (S + L) * D * (2 if method == "poe" else 1)
        macs += D * 256 + 256 * (2 * L + 2 * S) + dec
    return 6 * macs


def hbn_epoch_plans(n_steps, seed=0):
    from mopoe_b200 import data
    cohort = data.make_cohort()
    train = np.r_[0:2048, 2560:2560 + 512 + 256]
    has = np.stack([cohort["has_clinical"][train], cohort["has_rois"][train]])
    rng = np.random.RandomState(seed)
    plan = []
    while len(plan) < n_steps:
        plan += data.epoch_plan(has, 256, rng)
    return cohort, train, plan[:n_steps]


def cpu_train_sample(budget_s, method="joint_elbo"):
    """The reference training step (run_epochs.py:158-182: forward, ELBO, backward, Adam) through the oracle
    port on the HBN epoch plan, torch CPU with every host thread, for about `budget_s` seconds.
    -> (rows, seconds, steps, description)."""
    from oracle import mopoe_oracle as mo
    torch.set_num_threads(os.cpu_count())
    spec = mo.ModelSpec(**dict(HBN, method=method))
    params = mo.init_params(spec, seed=0)
    cohort, train, plan = hbn_epoch_plans(4000)
    xs = [torch.from_numpy(cohort["clinical"][train]), torch.from_numpy(cohort["rois"][train])]
    g = torch.Generator().manual_seed(0)
    opt = mo.Adam(params, lr=0.002)
    n_pass = 3 if method == "poe" else 1

    def one(i, params):
        mask, ix = plan[i]
        ix = torch.from_numpy(ix.astype(np.int64))
        batch = {n: xs[m][ix] for m, n in enumerate(spec.mod_names) if mask >> m & 1}
        eps = torch.randn(n_pass, len(ix), spec.eps_width, generator=g)
        out, gr, used = mo.elbo_and_grads(params, spec, batch, eps)
        return opt.step(params, gr, used), len(ix)
    for i in range(5):
        params, _ = one(i, params)
    rows, steps, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and 5 + steps < len(plan):
        params, n = one(5 + steps, params)
        rows += n; steps += 1
    dt = time.perf_counter() - t0
    return rows, dt, steps, ("%d fwd+bwd+Adam steps of the HBN epoch plan (batch 256, missing blocks allowed, %s) through the "
                             "torch-CPU port of run_epochs.train" % (steps, method))


def bench_train(device, peaks, steps=300, warmup=60):
    """BASELINE.json configs[1-2] and the training half of configs[4]: the fused fwd+bwd+Adam persistent kernel.
    HBN: `steps` consecutive batches of the MissingModalitySampler epoch plan (batch 256, missing blocks allowed) in
    ONE launch, for poe / moe / joint_elbo.  Stress: 4 modalities, 15 subsets, batch 65 536 (tensor-core kernel).
    Device-resident numbers are CUDA-event timed; e2e adds the H2D of the cohort blocks and batch plan and the D2H of
    the per-step scalar rows (second call: the first one pays allocator / module-load costs)."""
    import mopoe_b200
    from mopoe_b200 import _lib, engine
    out = {}
    cohort, train, plan = hbn_epoch_plans(steps + warmup)
    host = [torch.from_numpy(cohort["clinical"][train]).pin_memory(), torch.from_numpy(cohort["rois"][train]).pin_memory()]
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    impl_name = {0: "cuda-core persistent kernel", 1: "tcgen05 persistent kernel"}
    for method in ("joint_elbo", "moe", "poe"):
        spec = mopoe_b200.PathSpec(HBN["dims"], HBN["style_dims"], HBN["latent_dim"], method, HBN["mod_names"])
        flat = engine.pack_params(spec, engine.init_params(spec, seed=0), device)
        rows = sum(len(ix) for _, ix in plan[warmup:])
        offs = np.cumsum([0] + [len(ix) for _, ix in plan])
        index_h = torch.from_numpy(np.concatenate([ix for _, ix in plan]).astype(np.int32)).pin_memory()
        index = index_h.to(device)
        m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
        t_ = torch.zeros(4, dtype=torch.int32, device=device)
        ws = engine.Workspace()
        dev_data = [h.to(device, non_blocking=True) for h in host]
        blist = lambda lo, hi: [(len(plan[i][1]), plan[i][0], int(offs[i])) for i in range(lo, hi)]

        def launch(lo, hi, dd, idx):
            b = engine.make_batches(spec, blist(lo, hi), device)
            return engine.train_steps(spec, flat, dd, b, hi - lo, 256, 2, row_index=[idx, idx], seed=7,
                                      adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
        launch(0, warmup, dev_data, index)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bdesc = engine.make_batches(spec, blist(warmup, warmup + steps), device)
        torch.cuda.synchronize()
        e0.record()
        sc = engine.train_steps(spec, flat, dev_data, bdesc, steps, 256, 2, row_index=[index, index], seed=7,
                                adam_m=m_, adam_v=v_, adam_t=t_, lr=0.002, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        impl = _lib.lib().mopoe_train_last_impl()
        host_sc = torch.empty(steps, _lib.N_SCALARS).pin_memory()

        def e2e_once():
            dd = [h.to(device, non_blocking=True) for h in host]
            idx = index_h.to(device, non_blocking=True)
            host_sc.copy_(launch(warmup, warmup + steps, dd, idx), non_blocking=True)
            torch.cuda.synchronize()
        e2e_once()                                   # warm the path (allocations, first-call costs)
        t0 = time.perf_counter()
        e2e_once()
        e2e_s = time.perf_counter() - t0
        flop = model_flops_per_row(HBN, method) * rows / steps
        out[method] = {"samples_per_s": rows / (ms * 1e-3), "us_per_step": 1e3 * ms / steps, "steps_per_launch": steps,
                       "impl": impl_name[impl], "gpu_launches": 1,
                       "e2e": {"samples_per_s": rows / e2e_s, "h2d_bytes": sum(h.numel() * 4 for h in host) + index_h.numel() * 4,
                               "d2h_bytes": host_sc.numel() * 4},
                       "final_loss": float(host_sc[-1, 0]), "flop_per_step": flop,
                       "roofline": {"bound": "latency (grid barriers + dependent stages: 0.25 GFLOP and 0.67 MB of weights per step)",
                                    "achieved_tflops": flop / (ms * 1e-3 / steps) / 1e12, "peak_tflops": tf_peak,
                                    "frac": flop / (ms * 1e-3 / steps) / 1e12 / tf_peak,
                                    "launches_eliminated_per_step": "~1000 eager kernels + ~10 .item() syncs -> 1/%d launch" % steps}}
    # stress shape (configs[4], training half): batch 65 536, 4 modalities, 15 subsets
    try:
        for method in ("joint_elbo", "poe"):
            spec = mopoe_b200.PathSpec(STRESS["dims"], STRESS["style_dims"], 20, method, STRESS["mod_names"])
            flat = engine.pack_params(spec, engine.init_params(spec, seed=0), device)
            g = torch.Generator().manual_seed(0)
            n, k = 65536, 4
            dd = [torch.randn(n, d, generator=g).to(device) for d in spec.dims]
            idx = torch.arange(n, dtype=torch.int32, device=device)
            bdev = engine.make_batches(spec, [(n, 15, 0)] * k, device)
            m_, v_ = torch.zeros_like(flat), torch.zeros_like(flat)
            t_ = torch.zeros(4, dtype=torch.int32, device=device)
            ws = engine.Workspace()
            go = lambda: engine.train_steps(spec, flat, dd, bdev, k, n, 2, row_index=[idx] * 4, seed=7, adam_m=m_, adam_v=v_,
                                            adam_t=t_, lr=0.002, workspace=ws)
            go(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sc = go(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / k
            flop = model_flops_per_row(STRESS, method) * n
            out["stress_" + method] = {"samples_per_s": n / (ms * 1e-3), "ms_per_step": ms, "batch": n,
                                       "impl": impl_name[_lib.lib().mopoe_train_last_impl()], "final_loss": float(sc[-1, 0]),
                                       "flop_per_step": flop,
                                       "roofline": {"bound": "tensor", "achieved_tflops": flop / (ms * 1e-3) / 1e12, "peak_tflops": tf_peak,
                                                    "frac": flop / (ms * 1e-3) / 1e12 / tf_peak,
                                                    "note": "fp32-equivalent FLOPs; each contraction runs as 3 fp16 tensor-core passes (3xFP16 split)"}}
            del dd, ws
    except Exception as exc:
        out["stress_error"] = repr(exc)
    # stress shape (configs[4], DAA half): 4 modalities / 15 PoE subsets, 20 validations x 1 000 subjects x 7 scores x
    # 150 samples = 21 M avatars, 37.3 GB avatar tensor materialised in HBM (sized for 180 GB), pipelined tcgen05 kernel
    try:
        from mopoe_b200 import daa
        import ctypes as C
        spec = mopoe_b200.PathSpec(STRESS["dims"], STRESS["style_dims"], 20, "joint_elbo", STRESS["mod_names"])
        flat = engine.pack_params(spec, engine.init_params(spec, seed=0), device)
        g = torch.Generator().manual_seed(1)
        nv, ns, J, Mb = 20, 1000, 150, 1000
        xs = [torch.randn(nv, ns, d, generator=g).to(device) for d in spec.dims]
        ws = engine.Workspace()
        lib = _lib.lib()
        _lib.check(lib.mopoe_profile_enable(1))
        r = None
        go = lambda r: daa.daa_sweep(spec, flat, xs[0], xs[1], J, Mb, seed=3, others={2: xs[2], 3: xs[3]}, workspace=ws, out=r,
                                     base_mean="direct", want_betas=False)
        r = go(r); torch.cuda.synchronize()
        impl = int(lib.mopoe_daa_last_impl())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = go(r)
        e1.record(); torch.cuda.synchronize()
        kms = C.c_float()
        _lib.check(lib.mopoe_daa_last_kernel_ms(C.byref(kms)))
        ms = e0.elapsed_time(e1) / 3
        n_av = nv * ns * spec.dims[0] * J
        nbytes = n_av * (spec.dims[1] * 4 + 4)
        out["stress_daa"] = {"avatars_per_s": n_av / (ms * 1e-3), "ms_per_sweep": ms, "avatars": n_av, "n_subjects": ns, "n_validation": nv,
                             "avatar_tensor_gb": n_av * spec.dims[1] * 4 / 1e9, "impl": {2: "pipelined tcgen05 kernel", 1: "tcgen05 kernel", 0: "cuda-core kernel"}[impl],
                             "kernel_ms": kms.value, "kernel_gbs": nbytes / (kms.value * 1e-3) / 1e9,
                             "finite": bool(torch.isfinite(r.pvalues).all())}
        del r, xs, ws
        torch.cuda.empty_cache()
    except Exception as exc:
        out["stress_daa_error"] = repr(exc)
    return out


def run_ours(args, rank, world, local_rank):
    import mopoe_b200
    from mopoe_b200 import _lib, daa, engine
    import ctypes as C
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # one process per GPU: keep this rank's pinned host buffers on the GPU's NUMA node (device-to-host copies of
    # several ranks otherwise contend on the socket interconnect)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(visible.split(",")[local_rank]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else local_rank
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = None if os.environ.get("MOPOE_BENCH_NO_NUMA") else daa.bind_to_gpu_numa_node(phys)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    spec = mopoe_b200.PathSpec(HBN["dims"], HBN["style_dims"], HBN["latent_dim"], "joint_elbo", HBN["mod_names"])
    flat = engine.pack_params(spec, engine.init_params(spec, seed=0), device)
    n_val, N, J, Mb = DAA["n_validation"], DAA["n_subjects"], DAA["n_samples"], DAA["n_base"]
    C_, R = spec.dims[0], spec.dims[1]
    src_h, dst_h = draw_validation_batches(n_val, DAA["seed"], offset=rank)
    src_h, dst_h = src_h.pin_memory(), dst_h.pin_memory()
    src_d, dst_d = src_h.to(device), dst_h.to(device)
    ws = engine.Workspace()
    lib = _lib.lib()
    _lib.check(lib.mopoe_profile_enable(1))

    class Runner:
        """One shard of a sweep on this rank: sweep + exchange of the association tables, replayed from a CUDA graph."""

        def __init__(self, src, dst, val_begin, n_val_total, base_mean="direct", units=None):
            self.src, self.dst, self.val_begin, self.n_val_total, self.base_mean = src, dst, val_begin, n_val_total, base_mean
            self.units = units                          # daa.shard_units(...) of this rank: (validation, score) shard units
            self.ex, self.full, self.graph, self.replays = None, {}, None, 0
            self.exchange_note = "single GPU: nothing to exchange"
            if world > 1:
                ok = torch.zeros(1, device=device)
                if not os.environ.get("MOPOE_BENCH_NO_EXCHANGE"):
                    try:
                        self.ex = daa.TableExchange(n_val_total, C_, R, device, root=0)
                        ok += 1
                    except Exception as exc:
                        self.exchange_note = "NCCL all_gather (peer-memory exchange unavailable: %s)" % type(exc).__name__
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if ok.item() < 1:
                    self.ex = None
                    if self.exchange_note.startswith("single"):
                        self.exchange_note = "NCCL all_gather (peer-memory exchange disabled)"
                else:
                    self.exchange_note = ("peer-memory gather of the association tables to rank 0 (what daa_exp needs): NVLink P2P stores into "
                                          "rank 0's table + system-scope flags and acknowledgements, no NCCL launch; senders run at most "
                                          "two sweeps ahead of rank 0")
            self.r = self.sweep()
            self.gather()

        def sweep(self, src=None, dst=None, out=None):
            return daa.daa_sweep(spec, flat, self.src if src is None else src, self.dst if dst is None else dst, J, Mb,
                                 seed=DAA["seed"], val_begin=self.val_begin, n_val_total=self.n_val_total, workspace=ws, out=out,
                                 base_mean=self.base_mean,
                                 unit_begin=self.units["local_begin"] if self.units else None,
                                 unit_end=self.units["local_end"] if self.units else None)

        def gather(self, r=None):
            r = r or self.r
            if self.ex is not None and self.units:
                self.ex.exchange_units(r.coefs, r.pvalues, self.units["local_begin"], self.units["local_end"], self.units["unit_begin"])
            elif self.units and world > 1:
                u = self.units
                self.full["t"] = [daa.gather_tables(t.view(-1, R)[u["local_begin"]:u["local_end"]], self.n_val_total * C_).view(-1, C_, R)
                                  for t in (r.coefs, r.pvalues)]
            elif self.ex is not None:
                self.ex.exchange(r.coefs, r.pvalues, self.val_begin)
            elif world > 1:   # both tables in one coalesced NCCL launch, straight into reused full-size tensors
                self.full["t"] = daa.gather_tables_many([r.coefs, r.pvalues], self.n_val_total, outs=self.full.get("t"))

        def tables(self):
            if self.ex is not None:
                return self.ex.tables()
            return (self.full["t"][0], self.full["t"][1]) if world > 1 else (self.r.coefs, self.r.pvalues)

        def verify_exchange(self):
            """the pushed tables equal an NCCL all_gather of the same slices (unequal shards included)"""
            if self.ex is None:
                return True
            torch.cuda.synchronize()
            dist.barrier()
            cf, pv = self.tables()
            if self.units:
                u = self.units
                ref = [daa.gather_tables(t.view(-1, R)[u["local_begin"]:u["local_end"]], self.n_val_total * C_).view(-1, C_, R)
                       for t in (self.r.coefs, self.r.pvalues)]
            else:
                ref = [daa.gather_tables(t, self.n_val_total) for t in (self.r.coefs, self.r.pvalues)]
            torch.cuda.synchronize()
            ok = torch.ones(1, device=device)
            if rank == 0 and not (torch.equal(cf, ref[0]) and torch.equal(pv, ref[1])):     # the root holds the gathered tables
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            return bool(ok.item() > 0)

        def capture(self):
            """the whole step (8 kernels of the sweep on two streams + the table exchange) captured once and replayed;
            the replay is checked bit-for-bit against direct launches, direct launches are the fallback"""
            note = "direct launches"
            if os.environ.get("MOPOE_BENCH_NO_GRAPH"):
                return note
            try:
                own = (lambda t: t.view(-1, R)[self.units["local_begin"]:self.units["local_end"]]) if self.units else (lambda t: t)
                want = (own(self.r.coefs).clone(), own(self.r.pvalues).clone())
                _lib.check(lib.mopoe_profile_enable(0))
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self.r = self.sweep(out=self.r)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                calls = self.ex.calls if self.ex is not None else 0
                with torch.cuda.graph(g):
                    self.r = self.sweep(out=self.r)
                    if self.ex is not None:
                        self.gather()
                if self.ex is not None:
                    self.ex.calls = calls                 # captured, not executed
                self.r.coefs.zero_(); self.r.pvalues.zero_()
                g.replay()
                if self.ex is not None:
                    self.ex.calls += 1
                torch.cuda.synchronize()
                if torch.equal(own(self.r.coefs), want[0]) and torch.equal(own(self.r.pvalues), want[1]):
                    self.graph = g
                    note = "CUDA graph replay of the sweep%s (verified bit-identical to direct launches)" % (
                        " + table exchange" if self.ex is not None else "")
                else:
                    note = "direct launches (graph replay differed)"
            except Exception as exc:
                note = "direct launches (graph capture failed: %s)" % type(exc).__name__
                torch.cuda.synchronize()
            finally:
                _lib.check(lib.mopoe_profile_enable(1))
            return note

        def step(self):
            if self.graph is not None:
                self.graph.replay()
                if self.ex is not None:
                    self.ex.calls += 1
                else:
                    self.gather()
            else:
                self.r = self.sweep(out=self.r)
                self.gather()

        def timed(self, k):
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(k):
                self.step()
            a1.record()
            barrier()
            return a0.elapsed_time(a1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run = Runner(src_d, dst_d, rank * n_val, world * n_val)
    run_exchange_note = run.exchange_note
    for _ in range(max(0, args.warmup - 1)):
        run.step()
    exchange_ok = run.verify_exchange()
    graph_note = run.capture()
    r = run.r

    def sweep(src, dst, out=None):
        return run.sweep(src, dst, out)

    def gather(rr):
        run.gather(rr)
        return run.tables()

    def step():
        run.step()

    LAUNCH_NOTE[0] = graph_note
    LAUNCH_NOTE[1] = ("rank bound to the %d cores local to its GPU" % len(numa_cpus)) if numa_cpus else "none"
    for _ in range(2):
        step()
    # ---- value: device-resident inputs ----
    sampler = ClockSampler(local_rank)
    kernel_ms = []
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)   # (the clock sampler keeps running over the e2e regions: K sub-millisecond steps
                               # are shorter than one nvidia-smi sampling period)
    # dominant kernel alone (same launches, measured by the library's own event bracket)
    for _ in range(min(args.steps, 5)):
        r = sweep(src_d, dst_d, out=r)
        v = C.c_float()
        _lib.check(lib.mopoe_daa_last_kernel_ms(C.byref(v)))
        kernel_ms.append(v.value)
    # the same sweep with the M base passes drawn one by one (draw-for-draw what M reference forwards consume)
    drun = Runner(src_d, dst_d, rank * n_val, world * n_val, base_mean="draws") if world == 1 else None
    ms_draws = None
    if drun is not None:
        drun.ex = None
        drun.capture()
        for _ in range(2):
            drun.step()
        ms_draws = drun.timed(args.steps)
        del drun
    # ---- e2e: pinned host buffers in and out ----
    host_out = {k: torch.empty(getattr(r, k).shape, dtype=getattr(r, k).dtype).pin_memory()
                for k in ("avatars", "sampled_scores", "reconstructions", "betas", "coefs", "pvalues")}

    def e2e_step():
        s, d = src_h.to(device, non_blocking=True), dst_h.to(device, non_blocking=True)
        rr = sweep(s, d, out=r)
        cf, pv = gather(rr)
        for k in ("avatars", "sampled_scores", "reconstructions", "betas"):
            host_out[k].copy_(getattr(rr, k), non_blocking=True)
        host_out["coefs"].copy_(rr.coefs, non_blocking=True)
        host_out["pvalues"].copy_(rr.pvalues, non_blocking=True)

    e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # tables-only variant (avatar tensor stays in HBM)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        s, d = src_h.to(device, non_blocking=True), dst_h.to(device, non_blocking=True)
        rr = sweep(s, d, out=r)
        gather(rr)
        for k in ("sampled_scores", "reconstructions", "betas", "coefs", "pvalues"):
            host_out[k].copy_(getattr(rr, k), non_blocking=True)
    g1.record()
    barrier()
    ms_tab = g0.elapsed_time(g1)
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    d2h_tab = d2h - host_out["avatars"].numel() * 4
    # pinned device->host bandwidth of this rank (the bound of e2e: the reference also materialises the avatar tensor)
    probe = r.avatars.view(-1)[: 256 * 1024 * 1024 // 4]
    d2h_best = 0.0
    for _ in range(3):
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        host_out["avatars"].view(-1)[: probe.numel()].copy_(probe, non_blocking=True)
        h1.record()
        torch.cuda.synchronize()
        d2h_best = max(d2h_best, probe.numel() * 4 / (h0.elapsed_time(h1) * 1e-3) / 1e9)
    # ---- strong scaling (N > 1): the SAME 20 validations of configs[3] split over the ranks ----
    strong = None
    if world > 1:
        su = daa.shard_units(n_val, C_, rank, world)      # (validation, score) units, 140 for configs[3]
        sb, se = su["val_begin"], su["val_end"]
        s_src, s_dst = draw_validation_batches(n_val, DAA["seed"], offset=0)
        del run, r, host_out
        torch.cuda.empty_cache()
        srun = Runner(s_src[sb:se].to(device), s_dst[sb:se].to(device), sb, n_val, units=su)
        for _ in range(2):
            srun.step()
        s_ok = srun.verify_exchange()
        s_note = srun.capture()
        for _ in range(2):
            srun.step()
        ms_strong = srun.timed(args.steps)
        tms = torch.tensor([ms_strong], dtype=torch.float64, device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        shards = [daa.shard_units(n_val, C_, q, world) for q in range(world)]
        sizes = [q["unit_end"] - q["unit_begin"] for q in shards]
        strong = {"scaling": "strong", "n_validation_total": n_val, "shard_unit": "(validation, score)", "units_total": n_val * C_,
                  "units_per_rank": sizes, "validations_touched_per_rank": [q["val_end"] - q["val_begin"] for q in shards],
                  "ms_per_step": tms.item() / args.steps, "value": n_val * N * C_ * J * args.steps / (tms.item() * 1e-3), "unit": UNIT,
                  "balance_bound": n_val * C_ / (world * max(sizes)), "launch": s_note, "exchange": srun.exchange_note,
                  "exchange_verified_vs_nccl": bool(s_ok),
                  "note": "sharding unit = (validation, score) pair (SURVEY.md 8e): avatar tiles and statistics of the owned units only; "
                          "the base passes / encoder heads of a validation are run by every rank that shares it (they do not shrink "
                          "with the shard); the largest shard bounds the speed-up at balance_bound x N"}
    clocks = sampler.stop()
    times = torch.tensor([ms, ms_e2e, ms_tab, d2h_best], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times[:3], op=dist.ReduceOp.MAX)
        dist.all_reduce(times[3:], op=dist.ReduceOp.MIN)
    ms, ms_e2e, ms_tab, d2h_peak = times.tolist()
    avatars_per_step = world * n_val * N * C_ * J
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        k_ms = float(np.mean(kernel_ms))
        alg_bytes = n_val * N * C_ * J * (R * 4 + 4)        # avatar tile written + score read, per launch
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "daa_avatar_kernel_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        h2d = src_h.numel() * 4 + dst_h.numel() * 4
        line = {"metric": METRIC, "value": avatars_per_step * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world), "clocks": clocks,
                "e2e": {"value": avatars_per_step * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "d2h_peak_gbs_per_rank": d2h_peak, "d2h_achieved_gbs_per_rank": d2h * args.steps / (ms_e2e * 1e-3) / 1e9,
                        "frac_of_d2h_peak": d2h * args.steps / (ms_e2e * 1e-3) / 1e9 / max(d2h_peak, 1e-9),
                        "note": "every array daa_exp writes, incl. the 1.865 GB avatar tensor, copied to pinned host memory (buffers "
                                "allocated once); bound by the pinned device-to-host link measured in this run (256 MB copies, min over "
                                "ranks); with N ranks the copies share the host's memory system: the aggregate does not scale with N"},
                "e2e_tables_only": {"value": avatars_per_step * args.steps / (ms_tab * 1e-3), "unit": UNIT,
                                    "d2h_bytes_per_step": d2h_tab,
                                    "note": "avatar tensor left in HBM; scores, reconstructions, betas, coefs, p-values copied"},
                # per sweep: p1 + p2 (encoder heads, second stream), daa_base x 2 (noise | rest), operand prep, daa_avatar_pipe,
                # daa_beta_stats, daa_pvalue
                "gpu_launches": (8 if world == 1 else 10) * args.steps,   # + the two table-exchange kernels when N > 1
                "roofline": {"kernel": "daa_avatar_pipe_kernel", "bound": "hbm", "achieved": alg_bytes / (k_ms * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": traffic, "traffic_source": "ncu --set full capture committed under profiles/ (dram__bytes_read + write of "
                             "one launch of this kernel), not re-measured in this run", "kernel_ms": k_ms, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                             "algorithmic_bytes_per_launch": alg_bytes,
                             "flops": {"faithful_tflops": 331266.0 * n_val * N * C_ * J / (k_ms * 1e-3) / 1e12,
                                       "executed_tflops": 2 * 29800.0 * n_val * N * C_ * J / (k_ms * 1e-3) / 1e12,
                                       "note": "tcgen05 kind::f16 with the 3xFP16 split (fp32-level accuracy, fp32 accumulation in TMEM); "
                                               "faithful = 331 266 FLOP/avatar (reference recomputes both encoders), executed ~= 59.6 "
                                               "kFLOP/avatar (ROI encoder cached, rank-1 hidden update; x3 tensor-core passes not counted)"}}}
        line["config"]["exchange"] = run_exchange_note
        line["config"]["exchange_verified_vs_nccl"] = bool(exchange_ok)
        line["statistics_parity"] = "unpinned: statsmodels (stat_utils.make_regression) is absent; closed forms checked against scipy/numpy"
        # the bench sweeps random-init weights; the significance claim of north_star is carried by the GPU test below
        line["config"]["significance_parity"] = ("tests/test_gpu_parity.py::test_daa_full_sweep_trained_model_vs_oracle (trained model, this "
                                                 "workload, same Philox draws through the CPU oracle): 1 667 of 3 108 ROI-score pairs "
                                                 "significant at trust level 0.7, sets identical, significance margin 1.0e-4 log10 units, "
                                                 "avatars 1.5e-6 (recorded on B200, round 2)")
        if strong is not None:
            line["strong_scaling"] = strong
        if ms_draws is not None:
            line["value_base_draws"] = {"value": avatars_per_step * args.steps / (ms_draws * 1e-3), "unit": UNIT,
                                        "ms_per_step": ms_draws / args.steps,
                                        "note": "base_mean='draws': 43 M normals per sweep drawn and averaged (round-1 behaviour)"}
        sweep_bytes = alg_bytes
        line["roofline"]["whole_sweep"] = {"achieved": sweep_bytes / (ms / args.steps * 1e-3) / 1e9, "frac": sweep_bytes / (ms / args.steps * 1e-3) / 1e9 / hbm_peak,
                                           "note": "algorithmic bytes of the sweep over the WHOLE step time (all kernels, graph replay)"}
        if world == 1:
            os.sched_setaffinity(0, all_cpus)            # the CPU baseline uses every host core again
            n, dt, desc = cpu_daa_sample(12.0)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": desc}
            try:
                if not os.environ.get("MOPOE_BENCH_SKIP_TRAIN"):
                    tr = bench_train(device, peaks)
                    rows, dtt, st, tdesc = cpu_train_sample(8.0)
                    tr["cpu_baseline"] = {"value": rows / dtt, "unit": "samples/s", "us_per_step": 1e6 * dtt / max(1, st),
                                          "cores": os.cpu_count(), "kind": "port", "sample": tdesc}
                    tr["metric"] = "mopoe_train_samples_per_s"
                    tr["clocks"] = "see the line's `clocks` (the sampler spans the whole run)"
                    line["train"] = tr
            except Exception as exc:   # the headline line must still print
                line["train"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
