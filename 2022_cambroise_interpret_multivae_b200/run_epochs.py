"""Mirror of experiments/run_epochs.py on the B200 path.

  basic_routine_epoch(exp, model_idx, batch) -> {"results","log_probs","total_loss","klds"}
      same contract as run_epochs.py:73-135; ONE cooperative launch computes the forward, every
      loss term and the gradients; `total_loss.backward()` then only hands the gradients over.
  train / test / run_epochs
      same loop structure (run_epochs.py:138-256).  When the experiment carries a device-resident
      cohort (`exp.resident`, built by workflow.train_exp) an epoch is ONE persistent launch:
      fwd + bwd + Adam for every batch of the epoch plan, logging scalars written per step.
"""
import os

import numpy as np
import torch

from . import _lib, engine
from .data import epoch_plan
from .model import elbo_step


def _log_dicts(spec, sc, present_mask):
    keys = [k for k, _ in spec.subsets()]
    avail, _ = spec.mixture_subsets(present_mask)
    log_probs = {n: sc[_lib.S_NLL + m] for m, n in enumerate(spec.mod_names) if present_mask >> m & 1}
    klds = {keys[s]: sc[_lib.S_KLD_SUBSET + s] for s in avail}
    return log_probs, klds


def basic_routine_epoch(exp, model_idx, batch):
    model = exp.models
    if exp.flags.num_models > 1:
        model = model[model_idx]
    batch_d = batch[0]
    dev = next(model.parameters()).device
    for k in list(batch_d.keys()):
        batch_d[k] = batch_d[k].to(dev).float()                    # run_epochs.py:85-86
    sc, res, loss = elbo_step(model, batch_d, need_grad=torch.is_grad_enabled())
    log_probs, klds = _log_dicts(model.spec, sc, res.present_mask)
    return {"results": model._results(res, batch_d), "log_probs": log_probs, "total_loss": loss, "klds": klds}


class ScalarLog:
    """Stand-in for TBLogger (utils/TBLogger.py:84-101): keeps the per-step scalar rows that the
    fused kernel writes; `to_tensorboard` replays them into a SummaryWriter when one is available."""

    def __init__(self):
        self.train, self.test = [], []

    def add(self, phase, rows):
        getattr(self, phase).append(rows.detach().cpu().numpy().reshape(-1, _lib.N_SCALARS))

    def array(self, phase):
        rows = getattr(self, phase)
        return np.concatenate(rows) if rows else np.zeros((0, _lib.N_SCALARS), np.float32)

    def to_tensorboard(self, writer, spec):
        keys = [k for k, _ in spec.subsets()]
        for phase in ("train", "test"):
            for step, r in enumerate(self.array(phase)):
                mask = int(r[_lib.S_PRESENT])
                writer.add_scalars("%s/Loss" % phase, {"loss": float(r[0])}, step)
                writer.add_scalars("%s/LogProb" % phase, {n: float(r[_lib.S_NLL + m]) for m, n in enumerate(spec.mod_names) if mask >> m & 1}, step)
                avail, _ = spec.mixture_subsets(mask)
                writer.add_scalars("%s/KLD" % phase, {keys[s]: float(r[_lib.S_KLD_SUBSET + s]) for s in avail}, step)
                writer.add_scalars("%s/group_divergence" % phase, {"group_div": float(r[1])}, step)


def train(model_idx, epoch, exp, tb_logger):
    """One training epoch.  Fused path: the whole epoch plan in one persistent launch."""
    model = exp.models if exp.flags.num_models == 1 else exp.models[model_idx]
    model.train()
    spec, flags = model.spec, exp.flags
    res = exp.resident_of(model_idx)
    flat = model.flat_parameters()
    plan = epoch_plan(res["has_train"], flags.batch_size, exp.rng)
    if not flags.allow_missing_blocks:
        plan = [p for p in plan if p[0] == (1 << spec.n_mods) - 1]
    offs = np.cumsum([0] + [len(ix) for _, ix in plan])
    index = torch.from_numpy(np.concatenate([ix for _, ix in plan]).astype(np.int32)).to(flat.device)
    bdev = engine.make_batches(spec, [(len(ix), mask, int(offs[i])) for i, (mask, ix) in enumerate(plan)], flat.device)
    st = exp.adam_of(model_idx)
    sc = engine.train_steps(spec, flat, res["train"], bdev, len(plan), flags.batch_size, 2,
                            row_index=[index] * spec.n_mods, seed=exp.next_seed(), adam_m=st["m"], adam_v=st["v"],
                            adam_t=st["t"], lr=flags.initial_learning_rate, b1=flags.beta_1, b2=flags.beta_2,
                            workspace=model._ws)
    tb_logger.add("train", sc)
    return sc


def test(model_idx, epoch, exp, tb_logger):
    """Test epoch (run_epochs.py:187-219): losses only, no gradient, test batches in order."""
    model = exp.models if exp.flags.num_models == 1 else exp.models[model_idx]
    model.eval()
    spec, flags = model.spec, exp.flags
    res = exp.resident_of(model_idx)
    flat = model.flat_parameters()
    n = res["n_test"]
    full = (1 << spec.n_mods) - 1
    blist = [(min(flags.batch_size, n - o), full, o) for o in range(0, n, flags.batch_size)]
    index = torch.arange(n, dtype=torch.int32, device=flat.device)
    bdev = engine.make_batches(spec, blist, flat.device)
    sc = engine.train_steps(spec, flat, res["test"], bdev, len(blist), flags.batch_size, 0,
                            row_index=[index] * spec.n_mods, seed=exp.next_seed(), workspace=model._ws)
    tb_logger.add("test", sc)
    return sc


def run_epochs(exp):
    """run_epochs.py:222-256: flags.rar, epochs of train+test, checkpoints every 5 epochs."""
    flags = exp.flags
    os.makedirs(flags.dir_experiment_run, exist_ok=True)
    torch.save(flags, os.path.join(flags.dir_experiment_run, "flags.rar"))       # utils.py:120-121
    logs = []
    for model_idx in range(flags.num_models):
        tb_logger = ScalarLog()
        logs.append(tb_logger)
        for epoch in range(flags.start_epoch, flags.end_epoch):
            train(model_idx, epoch, exp, tb_logger)
            test(model_idx, epoch, exp, tb_logger)
            if (epoch + 1) % 5 == 0 or (epoch + 1) == flags.end_epoch:
                d = os.path.join(flags.dir_checkpoints, str(epoch).zfill(4))
                model = exp.models if flags.num_models == 1 else exp.models[model_idx]
                if flags.num_models > 1:
                    d = os.path.join(flags.dir_checkpoints, "model_%d" % model_idx, str(epoch).zfill(4))
                os.makedirs(d, exist_ok=True)
                model.save_networks()
                torch.save(model.state_dict(), os.path.join(d, flags.model_save))
        np.save(os.path.join(flags.dir_logs, "scalars_train_model%d.npy" % model_idx), tb_logger.array("train"))
        np.save(os.path.join(flags.dir_logs, "scalars_test_model%d.npy" % model_idx), tb_logger.array("test"))
    return logs
