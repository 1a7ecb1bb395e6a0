"""Mirror of the two hot-path workflows of experiments/workflow.py on the B200 path.

  train_exp(dataset, datasetdir, outdir, input_dims, ...)      workflow.py:41-182
  daa_exp(dataset, datasetdir, outdir, run, ...)               workflow.py:185-539

Same keyword arguments, same run-directory layout, same output files (names, dtypes, axis order):
  <outdir>/<dataset>_<YYYY_MM_DD_HH_MM>/flags.rar, checkpoints/<epoch:04d>/model, checkpoints/enc_*,
  <run>/daa/<params>/rois_digital_avatars.npy (float32 (n_val, n_subj, n_scores, n_samples, n_rois)),
  sampled_scores.npy, metadatas.npy, rois_reconstructions.npy, coefs.npy, pvalues.npy,
  all_coefs.npy, significant_rois.tsv.
What differs (documented in DESIGN.md): the cohort is standardised once and kept resident in HBM
(no DataLoader workers), an epoch is one persistent launch, the DAA statistics are computed on the
GPU in fp64 in closed form, and the train/test split is our own seeded split (the reference's
iterstrat-based fetcher is out of scope; only subjects with every block go to the test set, as in
multiblock_fetcher.py:102-119).  With torch.distributed initialised, daa_exp shards validations
over the ranks and all-gathers the result tables.
"""
import glob
import os
import time
from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch

from . import daa, run_epochs as re_
from .engine import Workspace
from .model import VAE

MODALITIES = ["clinical", "rois"]


class _Mod:
    def __init__(self, name):
        self.name = name


class Experiment:
    """What MultimodalExperiment provides to the hot path (experiment.py:64-91), with the cohort
    standardised once (StandardScaler on the train rows, experiment.py:146-166) and resident in HBM.
    `flags.num_models > 1` builds the reference's ensemble (experiment.py:196-235): k-fold splits of the subjects
    (no held-out test set; model i is tested on fold i), one model, scaler and optimiser state per fold; every
    per-model attribute is then a list, as in the reference."""

    def __init__(self, flags, device=None):
        self.flags = flags
        self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.modalities = {n: _Mod(n) for n in MODALITIES[: len(flags.input_dim)]}
        self.mod_names = list(self.modalities)
        self.num_models = int(getattr(flags, "num_models", 1) or 1)
        self.rng = np.random.RandomState(None if flags.data_seed == "defaults" else int(flags.data_seed))
        self._seed = int(self.rng.randint(1, 2 ** 31 - 1))
        self._load(flags.datasetdir)
        models = [VAE(flags, self.modalities).to(self.device) for _ in range(self.num_models)]   # experiment.py:123-130
        self.models = models[0] if self.num_models == 1 else models
        self.optimizers = None
        self.adam_state = None
        self.rec_weights = {n: 1.0 for n in self.mod_names}
        self.style_weights = {n: flags.beta_style for n in self.mod_names}

    def next_seed(self):
        self._seed += 1
        return self._seed

    def model_of(self, model_idx):
        return self.models if self.num_models == 1 else self.models[model_idx]

    def resident_of(self, model_idx):
        return self.resident if self.num_models == 1 else self.resident[model_idx]

    def adam_of(self, model_idx):
        return self.adam_state if self.num_models == 1 else self.adam_state[model_idx]

    def _load(self, datasetdir):
        meta = pd.read_table(os.path.join(datasetdir, "metadata.tsv"))
        subjects = meta["participant_id"].to_numpy()
        pos = {s: i for i, s in enumerate(subjects)}
        n = len(subjects)
        blocks, has = [], []
        for mod in self.mod_names:
            x = np.load(os.path.join(datasetdir, mod + "_data.npy"), mmap_mode="r")
            subj = np.load(os.path.join(datasetdir, mod + "_subjects.npy"), allow_pickle=True)
            full = np.zeros((n, x.shape[1]), np.float32)
            h = np.zeros(n, bool)
            idx = np.array([pos[s] for s in subj])
            full[idx] = np.asarray(x, np.float32)
            h[idx] = True
            blocks.append(full)
            has.append(h)
        has = np.stack(has)
        complete = np.flatnonzero(has.all(0))
        split = np.random.RandomState(42)                       # fetchers/hbn.py:20 seed
        perm = split.permutation(complete)
        if self.num_models == 1:
            n_test = int(round(0.2 * len(complete)))            # experiment.py:203 test_size
            folds = [np.sort(perm[:n_test])]
        else:                                                   # experiment.py:203-207: validation = num_models, test_size = 0
            folds = [np.sort(f) for f in np.array_split(perm, self.num_models)]
        self.metadata = meta
        self.scalers, self.train_idx, self.test_idx, resident = [], [], [], []
        dev = self.device
        for test_idx in folds:
            train_mask = np.ones(n, bool)
            train_mask[test_idx] = False
            if not self.flags.allow_missing_blocks:
                train_mask &= has.all(0)
            train_idx = np.flatnonzero(train_mask & has.any(0))
            scalers, scaled = [], []
            for m in range(len(blocks)):
                rows = blocks[m][train_idx][has[m][train_idx]]
                mean, std = rows.mean(0), rows.std(0)
                std[std == 0] = 1.0
                scalers.append((mean, std))
                scaled.append(((blocks[m] - mean) / std).astype(np.float32))
            self.scalers.append(scalers)
            self.train_idx.append(train_idx)
            self.test_idx.append(test_idx)
            resident.append({"train": [torch.from_numpy(b[train_idx]).to(dev) for b in scaled],
                             "test": [torch.from_numpy(b[test_idx]).to(dev) for b in scaled],
                             "has_train": has[:, train_idx], "n_test": len(test_idx)})
        if self.num_models == 1:
            self.scalers, self.train_idx, self.test_idx, self.resident = self.scalers[0], self.train_idx[0], self.test_idx[0], resident[0]
        else:
            self.resident = resident

    def set_optimizers(self):
        states, total = [], 0
        for i in range(self.num_models):
            flat = self.model_of(i).flat_parameters()
            states.append({"m": torch.zeros_like(flat), "v": torch.zeros_like(flat),
                           "t": torch.zeros(4, dtype=torch.int32, device=flat.device)})
            total += sum(p.numel() for p in self.model_of(i).parameters())
        self.adam_state = states[0] if self.num_models == 1 else states
        print("num parameters: %d" % total)

    @classmethod
    def get_experiment(cls, flags_file, checkpoints_dir, load_epoch=None):
        """experiment.py:93-121 (torch >= 2.6 needs weights_only=False to unpickle the namespace)."""
        flags = torch.load(flags_file, weights_only=False)
        if "num_models" not in vars(flags):
            flags.num_models = 1
        flags.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        exp = cls(flags)
        for model_idx in range(exp.num_models):
            pattern = (os.path.join(checkpoints_dir, "*", flags.model_save) if exp.num_models == 1 else
                       os.path.join(checkpoints_dir, "model_%d" % model_idx, "*", flags.model_save))
            cp_files = glob.glob(pattern)
            if len(cp_files) == 0:
                raise ValueError("You need first to train the model.")
            cp_files = sorted(cp_files, key=lambda p: int(p.split(os.sep)[-2]))
            cp_file = cp_files[-1]
            if load_epoch is not None:
                epochs = np.array([int(p.split(os.sep)[-2]) for p in cp_files])
                cp_file = cp_files[int(np.argmin(epochs >= load_epoch))]
            print(cp_file)
            exp.model_of(model_idx).load_state_dict(torch.load(cp_file, map_location=flags.device))
        return exp, flags


def _make_flags(dataset, datasetdir, outdir, input_dims, num_models, latent_dim, style_dim, data_seed,
                num_hidden_layer_encoder, num_hidden_layer_decoder, allow_missing_blocks,
                factorized_representation, likelihood, learning_rate, batch_size, num_epochs, eval_freq,
                eval_freq_fid, beta, data_multiplications, dropout_rate, initial_out_logvar, learn_output_scale,
                out_scale_per_subject, method, grad_scaling):
    flags = SimpleNamespace(   # workflow.py:98-121, hot-path fields + the ones downstream tools read
        dataset=dataset, datasetdir=datasetdir, num_models=num_models, allow_missing_blocks=allow_missing_blocks,
        batch_size=batch_size, beta=beta, beta_1=0.9, beta_2=0.999, beta_content=1.0, beta_style=1.0,
        calc_nll=False, calc_prd=False, class_dim=latent_dim, data_multiplications=data_multiplications,
        dir_experiment=outdir, div_weight=None, div_weight_uniform_content=None, end_epoch=num_epochs,
        eval_freq=eval_freq, eval_freq_fid=eval_freq_fid, factorized_representation=factorized_representation,
        initial_learning_rate=learning_rate, initial_out_logvar=initial_out_logvar, input_dim=list(input_dims),
        joint_elbo=False, kl_annealing=0, include_prior_expert=False, learn_output_scale=learn_output_scale,
        learn_output_sample_scale=out_scale_per_subject, likelihood=likelihood, load_saved=False, method=method,
        model_save="model", modality_jsd=False, modality_moe=False, modality_poe=False,
        num_hidden_layer_encoder=num_hidden_layer_encoder, num_hidden_layer_decoder=num_hidden_layer_decoder,
        dropout_rate=dropout_rate, poe_unimodal_elbos=True, start_epoch=0, style_dim=list(style_dim),
        data_seed=data_seed, grad_scaling=grad_scaling)
    flags.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if method == "poe":
        flags.modality_poe = True
    elif method == "moe":
        flags.modality_moe = True
    elif method == "joint_elbo":
        flags.joint_elbo = True
    elif method == "jsd":
        flags.modality_jsd = True
    else:
        print("Method not implemented...exit!")                  # workflow.py:134-136
        return None
    flags.num_mods = len(flags.input_dim)
    flags.div_weight_uniform_content = 1 / (flags.num_mods + 1)
    flags.alpha_modalities = [flags.div_weight_uniform_content] + [1 / (flags.num_mods + 1)] * flags.num_mods
    if not flags.factorized_representation:
        flags.style_dim = [0] * len(flags.style_dim)
    return flags


def _create_dir_structure(flags):
    """utils/filehandling.py:13-17,29-94: <outdir>/<dataset>_<YYYY_MM_DD_HH_MM>/{checkpoints,logs}."""
    name = flags.dataset + "_" + time.strftime("%Y_%m_%d_%H_%M")
    flags.str_experiment = name
    flags.dir_experiment_run = os.path.join(flags.dir_experiment, name)
    flags.dir_checkpoints = os.path.join(flags.dir_experiment_run, "checkpoints")
    flags.dir_logs = os.path.join(flags.dir_experiment_run, "logs")
    for d in (flags.dir_experiment_run, flags.dir_checkpoints, flags.dir_logs):
        os.makedirs(d, exist_ok=True)
    return flags


def train_exp(dataset, datasetdir, outdir, input_dims, num_models=1, latent_dim=20, style_dim=[3, 20],
              data_seed="defaults", num_hidden_layer_encoder=1, num_hidden_layer_decoder=0,
              allow_missing_blocks=True, factorized_representation=True, likelihood="normal",
              learning_rate=0.002, batch_size=256, num_epochs=1500, eval_freq=25, eval_freq_fid=100, beta=1.,
              data_multiplications=1, dropout_rate=0., initial_out_logvar=-3., learn_output_scale=True,
              out_scale_per_subject=False, method="joint_elbo", grad_scaling=False):
    """Train the model (workflow.py:41-182).  Returns the run name."""
    flags = _make_flags(dataset, datasetdir, outdir, input_dims, num_models, latent_dim, style_dim, data_seed,
                        num_hidden_layer_encoder, num_hidden_layer_decoder, allow_missing_blocks,
                        factorized_representation, likelihood, learning_rate, batch_size, num_epochs, eval_freq,
                        eval_freq_fid, beta, data_multiplications, dropout_rate, initial_out_logvar,
                        learn_output_scale, out_scale_per_subject, method, grad_scaling)
    if flags is None:
        return None
    _create_dir_structure(flags)
    exp = Experiment(flags)
    exp.set_optimizers()
    exp.logs = re_.run_epochs(exp)
    runs_file = os.path.join(flags.dir_experiment, "runs.tsv")     # workflow.py:155-182
    row = pd.DataFrame(dict(name=[flags.str_experiment], dataset=[flags.dataset],
                            out_scale_per_subject=[flags.learn_output_sample_scale],
                            n_hidden_layer_encoder=[flags.num_hidden_layer_encoder],
                            n_hidden_layer_decoder=[flags.num_hidden_layer_decoder],
                            allow_missing_blocks=[flags.allow_missing_blocks]))
    if os.path.exists(runs_file):
        row = pd.concat((pd.read_table(runs_file), row))
    row.to_csv(runs_file, index=False, sep="\t")
    return flags.str_experiment


def significant_votes(pvalues, trust_level, n_models=1, vote_prop=1):
    """workflow.py:517-525: Bonferroni threshold 0.05 / n_rois / n_scores, vote over the validations of each model
    (>= trust_level * n_validation), then over the models of an ensemble (>= vote_prop * n_models).
    pvalues (n_val, C, R) or (n_models, n_val, C, R) -> bool (C, R)."""
    p = np.asarray(pvalues)
    if n_models == 1:
        return daa.significant(p, trust_level)
    per_model = np.stack([daa.significant(p[i], trust_level) for i in range(n_models)])
    return per_model.sum(0) >= vote_prop * n_models


def daa_exp(dataset, datasetdir, outdir, run, sampling_strategy="likelihood", n_validation=5, n_samples=200,
            n_subjects=50, M=1000, trust_level=0.75, seed=1037, reg_method="hierarchical", sample_latents=True,
            vote_prop=1, materialize_avatars=True):
    """Digital avatars analysis (workflow.py:185-539).  Returns the results directory."""
    if sampling_strategy not in ("likelihood", "linear"):
        # "uniform" indexes the wrong axis of its (N, n_scores, n_samples) array in the reference (workflow.py:347-350
        # vs :411-412) and "gaussian" is named (:220-222) but never built
        raise NotImplementedError("sampling_strategy=%r is not on the B200 path (likelihood, linear)" % (sampling_strategy,))
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
    expdir = os.path.join(outdir, run)
    daadir = os.path.join(expdir, "daa")
    flags_file = os.path.join(expdir, "flags.rar")
    if not os.path.isfile(flags_file):
        raise ValueError("You need first to train the model.")
    exp, flags = Experiment.get_experiment(flags_file, os.path.join(expdir, "checkpoints"))
    clinical_names = np.load(os.path.join(datasetdir, "clinical_names.npy"), allow_pickle=True)
    rois_names = np.load(os.path.join(datasetdir, "rois_names.npy"), allow_pickle=True)
    n_scores, n_rois = len(clinical_names), len(rois_names)
    params = SimpleNamespace(n_validation=n_validation, n_subjects=n_subjects, M=M, n_samples=n_samples,
                             reg_method=reg_method, sampling=sampling_strategy, sample_latents=sample_latents, seed=seed)
    name = "_".join(["_".join([k, str(v)]) for k, v in params.__dict__.items()])     # workflow.py:261-262
    resdir = os.path.join(daadir, name)
    if rank == 0:
        os.makedirs(resdir, exist_ok=True)
    n_models = exp.num_models
    lead = (n_models,) if n_models > 1 else ()                   # workflow.py:264-288: leading model axis for ensembles
    n_test_min = min(exp.resident_of(i)["n_test"] for i in range(n_models))
    if n_subjects > n_test_min:
        raise ValueError("n_subjects=%d exceeds the %d test subjects with every block" % (n_subjects, n_test_min))
    if world > n_validation * n_scores:
        raise ValueError("world size %d exceeds the %d (validation, score) shard units" % (world, n_validation * n_scores))
    # draw n_validation batches of test subjects per model on the host (workflow.py:362-372): rank 0 draws, every
    # rank uses the same draws (with seed=None each rank would otherwise draw its own)
    if seed is None:
        seed_t = torch.zeros(1, dtype=torch.int64, device=exp.device)
        if rank == 0:
            seed_t[0] = int(np.random.randint(0, 2 ** 31 - 1))
        if world > 1:
            dist.broadcast(seed_t, 0)
        draw_seed = int(seed_t.item())
    else:
        draw_seed = int(seed)
    rs = np.random.RandomState(draw_seed)
    # the shard unit is the (validation, score) pair; a rank runs the base passes of every validation
    # it touches and the avatars / statistics of its own units only
    shards = [daa.shard_units(n_validation, n_scores, q, world) for q in range(world)]
    sh = shards[rank]
    begin, end = sh["val_begin"], sh["val_end"]
    unit_sizes = [q["unit_end"] - q["unit_begin"] for q in shards]
    # per-validation arrays (scores, reconstructions) are identical on every rank that shares the validation: the
    # owner of its first unit contributes them
    lead_of = lambda q: [v for v in range(q["val_begin"], q["val_end"]) if q["unit_begin"] <= v * n_scores < q["unit_end"]]
    lead_sizes = [len(lead_of(q)) for q in shards]
    all_draws, tabs = [], {k: [] for k in ("coefs", "pvalues", "betas", "scores", "recons")}
    if materialize_avatars:
        from numpy.lib.format import open_memmap
        da_file = os.path.join(resdir, "rois_digital_avatars.npy")
        if world > 1:
            dist.barrier()
        if rank == 0:
            mm = open_memmap(da_file, dtype="float32", mode="w+", shape=lead + (n_validation, n_subjects, n_scores, n_samples, n_rois))
            del mm
        if world > 1:
            dist.barrier()
    for model_idx in range(n_models):
        model = exp.model_of(model_idx)
        model.eval()
        flat = model.flat_parameters()
        res = exp.resident_of(model_idx)
        draws = np.stack([rs.permutation(res["n_test"])[:n_subjects] for _ in range(n_validation)])
        all_draws.append(draws)
        idx = torch.from_numpy(draws[begin:end]).to(flat.device)
        src = res["test"][0][idx]
        dst = res["test"][1][idx]
        given = None
        if sampling_strategy == "linear":         # workflow.py:337-346: the same ramp for every subject, between the
            complete = torch.from_numpy(res["has_train"].all(0)).to(flat.device)   # 5 % / 95 % quantiles of the train
            lo_hi = np.quantile(res["train"][0][complete].cpu().numpy(), [0.05, 0.95], 0)     # subjects with every block
            ramp = torch.from_numpy(np.linspace(lo_hi[0], lo_hi[1], n_samples).astype(np.float32)).to(flat.device)
            given = ramp[None, :, None, :].expand(end - begin, n_samples, n_subjects, n_scores).contiguous()
        r = daa.daa_sweep(model.spec, flat, src, dst, n_samples, M, sample_latents=sample_latents, reg_method=reg_method,
                          seed=draw_seed + 7919 * model_idx, val_begin=begin, n_val_total=n_validation,
                          materialize=materialize_avatars, workspace=model._ws, scores=given,
                          unit_begin=sh["local_begin"] if world > 1 else None, unit_end=sh["local_end"] if world > 1 else None)
        daa.check_status(model.spec, r)           # a device-side protocol error must not end up in result files
        if world > 1:
            lb, le = sh["local_begin"], sh["local_end"]
            mine = [v - begin for v in lead_of(sh)]
            rows = lambda t: t.reshape((t.shape[0] * t.shape[1],) + tuple(t.shape[2:]))[lb:le]
            unrows = lambda t: t.reshape((n_validation, n_scores) + tuple(t.shape[1:]))
            got = [unrows(daa.gather_rows(rows(t), unit_sizes)) for t in (r.coefs, r.pvalues, r.betas)]
            got += [daa.gather_rows(t[mine], lead_sizes) for t in (r.sampled_scores, r.reconstructions)]
        else:
            got = [r.coefs, r.pvalues, r.betas, r.sampled_scores, r.reconstructions]
        torch.cuda.synchronize()
        for k, t in zip(("coefs", "pvalues", "betas", "scores", "recons"), got):
            tabs[k].append(t.cpu().numpy())
        if materialize_avatars:
            mm = np.load(da_file, mmap_mode="r+")
            out = mm[model_idx] if n_models > 1 else mm
            for v in range(begin, end):                          # disjoint (validation, :, score range) slices per rank
                c0 = max(sh["unit_begin"], v * n_scores) - v * n_scores
                c1 = min(sh["unit_end"], (v + 1) * n_scores) - v * n_scores
                part = r.avatars[v - begin, :, c0:c1]
                host = torch.empty(part.shape, dtype=torch.float32).pin_memory()   # pinned: the copy runs at link speed
                host.copy_(part)
                out[v, :, c0:c1] = host.numpy()
            mm.flush()
            del mm, out, host
    stack = lambda k: np.stack(tabs[k]) if n_models > 1 else tabs[k][0]
    coefs, pvalues, betas = stack("coefs"), stack("pvalues"), stack("betas")
    if rank == 0:
        test_idx = [exp.test_idx] if n_models == 1 else exp.test_idx
        metas = [exp.metadata.iloc[ti].reset_index(drop=True) for ti in test_idx]
        meta_cols = list(metas[0].columns)
        metadatas = [np.stack([metas[i].iloc[d].to_numpy() for d in all_draws[i]]) for i in range(n_models)]
        np.save(os.path.join(resdir, "sampled_scores.npy"), stack("scores"))
        np.save(os.path.join(resdir, "metadatas.npy"), np.stack(metadatas) if n_models > 1 else metadatas[0])
        np.save(os.path.join(resdir, "rois_reconstructions.npy"), stack("recons"))
        np.save(os.path.join(resdir, "pvalues.npy"), pvalues)
        np.save(os.path.join(resdir, "coefs.npy"), coefs)
        if reg_method == "hierarchical":                        # workflow.py:476-505: per-subject betas
            cols = [str(n).replace("&", "_").replace("-", "_") for n in rois_names]
            pid, site = meta_cols.index("participant_id"), meta_cols.index("site")
            all_coefs = []
            for i in range(n_models):
                b = betas[i] if n_models > 1 else betas
                per_model = []
                for v in range(n_validation):
                    per_model.append([])
                    m = metas[i].iloc[all_draws[i][v]].to_numpy()[:, [pid, site]]
                    for c in range(n_scores):
                        df = pd.DataFrame(m, columns=["participant_id", "site"])
                        per_model[v].append(pd.concat([df, pd.DataFrame(b[v, c], columns=cols)], axis=1))
                all_coefs.append(per_model)
            np.save(os.path.join(resdir, "all_coefs.npy"), np.array(all_coefs if n_models > 1 else all_coefs[0], dtype=object),
                    allow_pickle=True)
        idx_sign = significant_votes(pvalues, trust_level, n_models, vote_prop)       # workflow.py:517-525
        rows = {"metric": [], "roi": [], "score": []}
        for i, score in enumerate(clinical_names):
            for nm in rois_names[np.where(idx_sign[i])]:
                roi, metric = nm.rsplit("_", 1)
                rows["score"].append(score); rows["metric"].append(metric); rows["roi"].append(roi)
        pd.DataFrame.from_dict(rows).to_csv(os.path.join(resdir, "significant_rois.tsv"), sep="\t", index=False)
    if world > 1:
        dist.barrier()
    return resdir


def rsa_exp(dataset, datasetdir, outdir, run, n_validation=1, n_subjects=301, sample_latents=False, seed=None):
    """Representational similarity analysis of the latent spaces (workflow.py:656-820): per model, validation and
    latent block (joint, the clinical_rois subset posterior, the two style posteriors) the Euclidean dissimilarity
    matrix of `n_subjects` test subjects is compared with the dissimilarity matrix of every clinical score and
    covariate (age, sex, site[, fsiq]) by Kendall's tau-b.  Matrices and the P^2 pair counts run on the GPU
    (rsa.py / csrc/mopoe_rsa.cu).  `seed` (ours): subject draws; the reference shuffles with the global torch RNG.
    Returns the rsa directory; writes kendalltau_stats.npy (n_models, 4, n_validation, n_scores + n_cov, 2),
    latent_dissimilarity.npy, scores_dissimilarity.npy and kendalltau_<latent>.tsv like the reference."""
    from . import rsa
    expdir = os.path.join(outdir, run)
    rsadir = os.path.join(expdir, "rsa")
    flags_file = os.path.join(expdir, "flags.rar")
    if not os.path.isfile(flags_file):
        raise ValueError("You need first to train the model.")
    exp, flags = Experiment.get_experiment(flags_file, os.path.join(expdir, "checkpoints"))
    os.makedirs(rsadir, exist_ok=True)
    clinical_names = np.load(os.path.join(datasetdir, "clinical_names.npy"), allow_pickle=True)
    cov_names = ["age", "sex", "site"] + (["fsiq"] if dataset == "euaims" else [])
    categorical_covs = ["sex", "site"]
    latent_names = ["joint", "clinical_rois", "clinical_style", "rois_style"]
    n_models, n_scores = exp.num_models, len(clinical_names)
    kendalltaus = np.zeros((n_models, len(latent_names), n_validation, n_scores + len(cov_names), 2))
    latent_dis, scores_dis = [], []
    rs = np.random.RandomState(seed)
    ws = Workspace()
    for model_idx in range(n_models):
        model = exp.model_of(model_idx)
        model.eval()
        res = exp.resident_of(model_idx)
        test_idx = exp.test_idx if n_models == 1 else exp.test_idx[model_idx]
        meta = exp.metadata.iloc[test_idx].reset_index(drop=True)
        latent_dis.append([])
        scores_dis.append([])
        for val_idx in range(n_validation):
            take = rs.permutation(res["n_test"])[:n_subjects]              # workflow.py:733-741: one shuffled batch
            idx = torch.from_numpy(take).to(exp.device)
            data = {"clinical": res["test"][0][idx], "rois": res["test"][1][idx]}
            # the reference matrices do not depend on the latent block: built once per validation
            refs = [rsa.vec2cmat(data["clinical"][:, c].contiguous()) for c in range(n_scores)]
            refs += [rsa.vec2cmat(meta[name].to_numpy()[take], categorical=name in categorical_covs) for name in cov_names]
            refs = torch.stack(refs)
            for latent_idx, latent_name in enumerate(latent_names):
                latents = model(data, sample_latents=sample_latents)["latents"]       # workflow.py:751-758
                if latent_name == "joint":
                    latents = latents["joint"]
                elif "style" in latent_name:
                    latents = latents["modalities"][latent_name]
                else:
                    latents = latents["subsets"][latent_name]
                latents = model.reparameterize(latents[0], latents[1]) if sample_latents and latents[0] is not None else latents[0]
                if latents is not None:
                    cmat = rsa.data2cmat(latents)
                    taus, pvals = rsa.fit_rsa(cmat, refs, workspace=ws)
                    kendalltaus[model_idx, latent_idx, val_idx, :, 0] = taus
                    kendalltaus[model_idx, latent_idx, val_idx, :, 1] = pvals
                    latent_dis[model_idx].append(cmat.cpu().numpy())
                scores_dis[model_idx].append(refs.cpu().numpy())
    np.save(os.path.join(rsadir, "kendalltau_stats.npy"), kendalltaus)
    np.save(os.path.join(rsadir, "latent_dissimilarity.npy"), np.asarray(latent_dis))
    np.save(os.path.join(rsadir, "scores_dissimilarity.npy"), np.asarray(scores_dis))
    for latent_idx, latent_name in enumerate(latent_names):                          # workflow.py:797-815
        names = list(clinical_names) + cov_names
        k = kendalltaus[:, latent_idx]
        df = pd.DataFrame({"score": names, "pval": k[..., 1].mean((0, 1)), "pval_std": k[..., 1].std((0, 1)),
                           "r": k[..., 0].mean((0, 1)), "r_std": k[..., 0].std((0, 1))})
        df.to_csv(os.path.join(rsadir, "kendalltau_%s.tsv" % latent_name), sep="\t", index=False)
    return rsadir
