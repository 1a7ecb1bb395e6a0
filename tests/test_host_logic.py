"""CPU tests of the host side: C-ABI exports, model description, subset order, selection
boundaries, parameter layout, epoch planner, sharding and the world_size-2 gather (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import mopoe_b200
from mopoe_b200 import _lib, daa, data
from oracle import mopoe_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mopoe_b200.h")).read()
    declared = set(re.findall(r"\b(mopoe_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert set(_lib.EXPORTED) == declared
    assert not hasattr(L, "mopoe_umma_selftest")          # test code stays out of the product library
    T = ctypes.CDLL(_lib.SELFTEST_LIB_PATH)
    thdr = open(os.path.join(ROOT, "include", "mopoe_b200_selftest.h")).read()
    for name in set(re.findall(r"\b(mopoe_[a-z_0-9]+)\s*\(", thdr)):
        assert hasattr(T, name), name


def test_package_init_params_equal_the_oracle_init():
    """bench.py's product arm initialises weights inside the package (nothing from oracle/ on that path)."""
    from mopoe_b200 import engine
    from oracle import cases, mopoe_oracle as mo
    for base in (cases.HBN, cases.STRESS):
        spec = mopoe_b200.PathSpec(base["dims"], base["style_dims"], base["latent_dim"], "joint_elbo", base["mod_names"])
        got, want = engine.init_params(spec, seed=3), mo.init_params(mo.ModelSpec(**base), seed=3)
        assert list(got) == list(want)
        assert all(torch.equal(got[k], want[k]) for k in want)
    src = open(os.path.join(ROOT, "bench.py")).read()
    ours = src[src.index("def bench_train"):src.index("def main")]
    assert "from oracle" not in ours and "import oracle" not in ours, "product arm of bench.py imports oracle/"


def test_no_device_is_a_loud_error():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mopoe_b200 import engine
    spec = mopoe_b200.PathSpec([7, 444], [3, 20])
    with pytest.raises(_lib.MopoeError):
        engine.forward(spec, torch.zeros(spec.layout.total), {"clinical": torch.zeros(4, 7)})
    assert _lib.lib().mopoe_philox_normal(1, 1, 0, 4, None, None) == -2      # MOPOE_ENODEV
    assert b"no CUDA device" in _lib.lib().mopoe_last_error()


@pytest.mark.parametrize("kw", [dict(num_hidden_layer_decoder=5), dict(num_hidden_layer_encoder=-1), dict(latent_dim=64)])
def test_unsupported_configurations_are_rejected(kw):
    with pytest.raises(_lib.MopoeError):
        mopoe_b200.PathSpec([7, 444], [3, 20], **kw)


def test_unsupported_method_and_likelihood():
    with pytest.raises(NotImplementedError):
        mopoe_b200.PathSpec([7, 444], [3, 20], method="mvae")
    with pytest.raises(NotImplementedError):
        mopoe_b200.PathSpec([7, 444], [3, 20], likelihood="bernoulli")


def test_layered_architectures_keep_the_reference_state_dict_names():
    """networks.py:9-28,44-64 with hidden-layer counts other than (1, 0) and learn_output_sample_scale: parameter names,
    shapes and order equal the oracle's (pinned to the reference's state_dict by load_state_dict(strict=True) in
    oracle/make_golden.py), blocks are disjoint."""
    from oracle import mopoe_oracle as mo
    for kw, okw in ((dict(num_hidden_layer_encoder=0, num_hidden_layer_decoder=2, learn_output_sample_scale=True),
                     dict(n_hidden_enc=0, n_hidden_dec=2, sample_scale=True)),
                    (dict(num_hidden_layer_encoder=3), dict(n_hidden_enc=3)),
                    (dict(likelihood="laplace"), dict(likelihood="laplace"))):
        spec = mopoe_b200.PathSpec([7, 444], [3, 20], **kw)
        assert spec.layered
        want = mo.param_shapes(mo.ModelSpec(**okw))
        got = spec.param_slices()
        assert list(got) == list(want)
        assert all(tuple(got[k][1]) == tuple(want[k]) for k in want)
        spans = sorted((off, off + int(np.prod(shape))) for off, shape in got.values())
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= spec.layout.total
    assert not mopoe_b200.PathSpec([7, 444], [3, 20]).layered


def test_jsd_mixture_has_the_prior_component():
    """BaseMMVae.py:217-223: unimodal experts + the prior N(0, I); the last row range of the selection is the prior's."""
    spec = mopoe_b200.PathSpec([7, 444], [3, 20], method="jsd")
    assert spec.mixture_subsets(0b11)[1] == [0, 1]
    b = spec.batch_desc(256, 0b11)
    assert b.n_mix == 3 and list(b.joint_bounds[:4]) == [0, 85, 170, 256]
    b1 = spec.batch_desc(50, 0b10)
    assert b1.n_mix == 2 and list(b1.joint_bounds[:3]) == [0, 25, 50]


@pytest.mark.parametrize("names", [["clinical", "rois"], ["clinical", "rois", "modc", "modd"], ["zeta", "alpha", "mid"]])
def test_subset_table_matches_oracle_order(names):
    dims, style = [5 + i for i in range(len(names))], [2] * len(names)
    spec = mopoe_b200.PathSpec(dims, style, mod_names=names)
    ospec = mo.ModelSpec(dims=dims, style_dims=style, mod_names=names)
    assert spec.subsets() == ospec.subsets()
    assert len(spec.subsets()) == 2 ** len(names) - 1


def test_selection_bounds_are_the_reference_expression():
    for n, k in [(256, 3), (50, 3), (37, 2), (65536, 15), (7, 3), (1, 1)]:
        assert mopoe_b200.selection_bounds(n, k) == mo.selection_bounds(n, k)
    b = mopoe_b200.PathSpec([7, 444], [3, 20]).batch_desc(50, 3)
    assert list(b.joint_bounds)[:4] == [0, 16, 32, 50] and b.n_mix == 3


def test_param_layout_is_disjoint_and_named_like_the_reference():
    spec = mopoe_b200.PathSpec([7, 444], [3, 20])
    ospec = mo.ModelSpec()
    sl = spec.param_slices()
    assert {k: v[1] for k, v in sl.items()} == mo.param_shapes(ospec)
    spans = sorted((off, off + int(np.prod(shape))) for off, shape in sl.values())
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= spec.layout.total
    assert sum(int(np.prod(s)) for _, s in sl.values()) == 167173      # factorised default


def test_epoch_plan_is_homogeneous_and_complete():
    c = data.make_cohort(n_both=300, n_clinical_only=70, n_rois_only=40)
    has = np.stack([c["has_clinical"], c["has_rois"]])
    plan = data.epoch_plan(has, 64, np.random.RandomState(0))
    seen = np.concatenate([ix for _, ix in plan])
    assert sorted(seen.tolist()) == list(range(410))
    sizes = [len(ix) for _, ix in plan]
    first_incomplete = next(i for i, s in enumerate(sizes) if s < 64)
    assert all(s == 64 for s in sizes[:first_incomplete]) and all(s < 64 for s in sizes[first_incomplete:])
    for mask, ix in plan:
        assert np.all(has[0][ix] == bool(mask & 1)) and np.all(has[1][ix] == bool(mask & 2))


def test_shard_validations_partition():
    for n, w in [(20, 1), (20, 8), (5, 2), (3, 4)]:
        spans = [daa.shard_validations(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


_GLOO = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import mopoe_b200
from mopoe_b200 import daa
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
out = None
for n_val in (5, 6, 6):          # unequal shards (padded path), equal shards (one collective, reused output)
    b, e = daa.shard_validations(n_val, rank, world)
    local = torch.arange(b, e, dtype=torch.float64).view(-1, 1, 1).expand(-1, 7, 11).contiguous() + 0.5
    full = daa.gather_tables(local, n_val, out=out)
    want = torch.arange(n_val, dtype=torch.float64).view(-1, 1, 1).expand(-1, 7, 11) + 0.5
    assert full.shape == (n_val, 7, 11) and torch.equal(full, want), (rank, full[:, 0, 0])
    if out is not None and tuple(out.shape) == (n_val, 7, 11):
        assert full.data_ptr() == out.data_ptr()          # reused, no new allocation
    out = full
    a, b = daa.gather_tables_many([local, 2 * local], n_val)
    assert torch.equal(a, want) and torch.equal(b, 2 * want)
# (validation, score) shard units as daa_exp gathers them: unit rows + the per-validation arrays of the lead rank
for n_val, C in ((3, 7), (1, 3), (4, 1)):
    shards = [daa.shard_units(n_val, C, q, world) for q in range(world)]
    sh = shards[rank]
    sizes = [q["unit_end"] - q["unit_begin"] for q in shards]
    table = torch.arange(n_val * C * 5, dtype=torch.float64).view(n_val, C, 5)
    local = table[sh["val_begin"]:sh["val_end"]].reshape(-1, 5)[sh["local_begin"]:sh["local_end"]].clone()
    full = daa.gather_rows(local, sizes).view(n_val, C, 5)
    assert torch.equal(full, table), (rank, n_val, C)
    lead = lambda q: [v for v in range(q["val_begin"], q["val_end"]) if q["unit_begin"] <= v * C < q["unit_end"]]
    per_val = torch.arange(n_val * 3, dtype=torch.float32).view(n_val, 3)
    got = daa.gather_rows(per_val[lead(sh)], [len(lead(q)) for q in shards])
    assert torch.equal(got, per_val), (rank, n_val, C)
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("n_val,C,world", [(20, 7, 8), (20, 7, 3), (2, 7, 8), (1, 3, 8), (5, 1, 2)])
def test_shard_units_partition_and_balance(n_val, C, world):
    """(validation, score) units, contiguous block split; at most one unit of imbalance."""
    from mopoe_b200 import daa
    seen, sizes = [], []
    for r in range(world):
        sh = daa.shard_units(n_val, C, r, world)
        n = sh["unit_end"] - sh["unit_begin"]
        sizes.append(n)
        seen += list(range(sh["unit_begin"], sh["unit_end"]))
        if n:
            assert sh["val_begin"] * C <= sh["unit_begin"] and sh["unit_end"] <= sh["val_end"] * C
            assert sh["val_begin"] == sh["unit_begin"] // C and sh["val_end"] == (sh["unit_end"] - 1) // C + 1
            assert sh["local_begin"] == sh["unit_begin"] - sh["val_begin"] * C
            assert sh["local_end"] - sh["local_begin"] == n
    assert seen == list(range(n_val * C))
    assert max(sizes) - min(sizes) <= 1
    if (n_val, C, world) == (20, 7, 8):
        assert sorted(sizes) == [17] * 4 + [18] * 4          # 97 % balance (whole validations: 3 / 2 = 83 %)


def test_gather_tables_world_size_2_gloo(tmp_path):
    script = tmp_path / "g.py"
    script.write_text(_GLOO % ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_epoch_plan_equals_the_reference_sampler_golden():
    """data.epoch_plan reproduces MissingModalitySampler.__iter__ (dataset.py:295-354) batch for batch under the
    same numpy seed: golden plans from the unmodified reference class (oracle/make_golden_sampler.py)."""
    import json
    import os
    from mopoe_b200 import data
    from oracle.make_golden_sampler import has_matrix
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_epoch_plans.json")))
    assert len(gold) >= 4
    for g in gold:
        case = g["case"]
        plan = data.epoch_plan(has_matrix(case), case["batch_size"], np.random.RandomState(case["seed"]))
        assert [ix.tolist() for _, ix in plan] == g["plan"], case



def test_reference_rng_order_matches_the_reference_consumption():
    """model._draw issues torch.randn in the shapes / order the reference consumes the generator
    (BaseMMVae.py:143-159, run_epochs.py:108-118): the same seed gives the tensors the oracle's
    reference_eps_list expects, for joint_elbo and poe, with a missing block."""
    from types import SimpleNamespace
    from mopoe_b200.model import VAE
    from oracle import mopoe_oracle as mo
    for method, keys in (("joint_elbo", ["clinical", "rois"]), ("poe", ["clinical", "rois"]), ("poe", ["rois"])):
        flags = SimpleNamespace(input_dim=[7, 444], style_dim=[3, 20], class_dim=20, factorized_representation=True,
                                modality_poe=method == "poe", modality_moe=False, modality_jsd=False, joint_elbo=method == "joint_elbo",
                                learn_output_scale=True, learn_output_sample_scale=False, beta=1.0, beta_style=1.0, beta_content=1.0,
                                num_hidden_layer_encoder=1, num_hidden_layer_decoder=0, likelihood="normal", initial_out_logvar=-3.0,
                                dropout_rate=0.0, num_models=1, dir_checkpoints="")
        model = VAE(flags, {"clinical": None, "rois": None})
        spec = model.spec
        N = 9
        torch.manual_seed(123)
        eps = model._draw(spec.n_pass, N, torch.device("cpu"), keys)
        ospec = mo.ModelSpec(method=method)
        present = [spec.mod_names.index(k) for k in keys]
        want_list = mo.reference_eps_list(ospec, present, eps)
        torch.manual_seed(123)
        for want in want_list:                       # the reference: one randn_like per reparameterize call
            got = torch.randn(want.shape)
            assert torch.equal(got, want)


def test_significant_votes_over_validations_and_models():
    """workflow.py:517-525: per-model vote over validations, then vote_prop over the models of an ensemble."""
    from mopoe_b200 import workflow
    rng = np.random.default_rng(0)
    C_, R = 3, 5
    thr = 0.05 / R / C_
    p = rng.uniform(0, 2 * thr, size=(4, 6, C_, R))          # 4 models x 6 validations
    per_model = (p < thr).sum(1) >= 0.5 * 6
    for vote in (0.25, 0.5, 1):
        want = per_model.sum(0) >= vote * 4
        assert np.array_equal(workflow.significant_votes(p, 0.5, 4, vote), want)
    assert np.array_equal(workflow.significant_votes(p[0], 0.5), per_model[0])
