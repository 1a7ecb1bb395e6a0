"""Digital Avatars Analysis on the GPU: the device half of workflow.daa_exp
(experiments/workflow.py:361-537) behind one call, plus its multi-GPU sharding.

`daa_sweep` runs the avatar generation AND the association statistics for a block of validations;
`shard_validations` / `gather_tables` implement SURVEY.md 8e: validations are independent given the
weights, so they are split over ranks with no data-path collective and only the final
(n_val, n_scores, n_rois) fp64 tables are gathered."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import Workspace, _f32, _ptr, _require_cuda, _stream
from .spec import PathSpec

REG_METHODS = {"hierarchical": 0, "fixed": 1}


class DaaResult:
    pass


def daa_sweep(spec: PathSpec, flat_params, src, dst, n_samples, n_base, *, src_mod=0, dst_mod=1,
              sample_latents=True, reg_method="hierarchical", seed=1037, val_begin=0, n_val_total=None,
              eps_base=None, eps_score=None, eps_av=None, materialize=True, want_betas=True,
              others=None, workspace=None, out=None, base_mean="draws", unit_begin=None, unit_end=None, scores=None):
    """src: (n_val, N, C) drawn test batches of the perturbed modality, dst: (n_val, N, R).
    others: optional {modality index: (n_val, N, D_m)} for models with more than two modalities.
    base_mean: how the mean over the M stochastic reconstructions (workflow.py:388-398) gets its noise --
      "draws"  average M drawn noise rows, draw for draw what M reference forwards consume (default);
      "direct" draw the mean row itself, N(0, 1/M): the default decoders are affine in z, so the mean of the M
               decodes is the decode of mu + sd * eps_mean.  Same distribution of every output, 1/M of the
               draws (in-kernel generator only).
    scores: (n_val, n_samples, N, C) artificial score values used as they are instead of draws around the base
      reconstruction (sampling_strategy "linear", workflow.py:337-346); exclusive with eps_score.
    unit_begin, unit_end: the OWNED (validation, score) units of this call, u = v_local * C + score (SURVEY.md 8e,
      `shard_units`): only their rows of coefs / pvalues / betas and their avatars are produced, the other scores of
      a shared first / last validation are another rank's.  Default: every unit.
    Returns a DaaResult with CUDA tensors avatars (or None), sampled_scores, reconstructions,
    betas (or None), coefs, pvalues."""
    if spec.n_hidden_enc != 1 or spec.n_hidden_dec != 0 or spec.learn_output_sample_scale:
        if base_mean != "draws" or others is not None:
            raise NotImplementedError("layered architectures: base_mean='draws', two modalities")
        # (a unit range is honoured trivially: the layered sweep computes whole validations, the owned rows are among them)
        return daa_sweep_layered(spec, flat_params, src, dst, n_samples, n_base, src_mod=src_mod, dst_mod=dst_mod,
                                 sample_latents=sample_latents, reg_method=reg_method, seed=seed, val_begin=val_begin,
                                 eps_base=eps_base, eps_score=eps_score, eps_av=eps_av, scores=scores,
                                 materialize=materialize, workspace=workspace)
    if base_mean not in ("draws", "direct"):
        raise ValueError("base_mean=%r (draws, direct)" % (base_mean,))
    if base_mean == "direct" and eps_base is not None:
        raise ValueError("base_mean='direct' uses the in-kernel generator: do not inject eps_base")
    if reg_method not in REG_METHODS:
        raise NotImplementedError("reg_method=%r is not on the B200 path (hierarchical, fixed)" % (reg_method,))
    _require_cuda(flat_params, "parameters")
    if n_samples < 128:
        import warnings
        warnings.warn("n_samples=%d < 128: the tcgen05 avatar kernels tile 128 samples per series, this sweep runs on the "
                      "CUDA-core avatar kernel (about 7x slower per avatar)" % n_samples, stacklevel=2)
    device = flat_params.device
    src, dst = _f32(src), _f32(dst)
    _require_cuda(src, "src")
    _require_cuda(dst, "dst")
    n_val, N, Cc = src.shape
    R = dst.shape[2]
    assert Cc == spec.dims[src_mod] and R == spec.dims[dst_mod] and dst.shape[:2] == (n_val, N)
    xs = [None] * spec.n_mods
    xs[src_mod], xs[dst_mod] = src, dst
    for m in range(spec.n_mods):
        if xs[m] is None:
            if others is None or m not in others:
                raise ValueError("DAA needs every modality present (missing %s)" % spec.mod_names[m])
            xs[m] = _f32(others[m])
    E = spec.eps_width
    if eps_base is not None:
        eps_base = _f32(eps_base); assert eps_base.shape == (n_val, n_base, N, E)
    if scores is not None:
        if eps_score is not None:
            raise ValueError("scores and eps_score are exclusive")
        eps_score = scores
    if eps_score is not None:
        eps_score = _f32(eps_score); assert eps_score.shape == (n_val, n_samples, N, Cc)
        _require_cuda(eps_score, "scores" if scores is not None else "eps_score")
    if eps_av is not None:
        eps_av = _f32(eps_av); assert eps_av.shape == (n_val, n_samples, Cc, N, E)
    q = _lib.DaaDesc(n_val=n_val, val_begin=val_begin, n_val_total=n_val_total or n_val, n_subjects=N,
                     n_samples=n_samples, n_base=n_base, src_mod=src_mod, dst_mod=dst_mod,
                     sample_latents=int(bool(sample_latents)), reg_method=REG_METHODS[reg_method],
                     base_mode=1 if base_mean == "direct" else 0,
                     unit_begin=0 if unit_end is None else int(unit_begin or 0), unit_end=0 if unit_end is None else int(unit_end),
                     score_mode=0 if scores is None else 1)
    bd = spec.batch_desc(N, (1 << spec.n_mods) - 1)
    lib = _lib.lib()
    nbytes = lib.mopoe_daa_workspace_bytes(C.byref(spec.desc), C.byref(q))
    if nbytes < 0:
        _lib.check(int(nbytes))
    ws = (workspace or Workspace()).get(nbytes, device)
    r = out or DaaResult()
    if out is None:
        f = lambda dt, *s: torch.empty(*s, dtype=dt, device=device)
        r.avatars = f(torch.float32, n_val, N, Cc, n_samples, R) if materialize else None
        r.sampled_scores = f(torch.float32, n_val, N, n_samples, Cc)
        r.reconstructions = f(torch.float32, n_val, N, R)
        r.betas = f(torch.float64, n_val, Cc, N, R) if want_betas else None
        r.coefs = f(torch.float64, n_val, Cc, R)
        r.pvalues = f(torch.float64, n_val, Cc, R)
    xp = (C.c_void_p * _lib.MAX_MODS)(*[_ptr(x).value for x in xs] + [None] * (_lib.MAX_MODS - spec.n_mods))
    _lib.check(lib.mopoe_daa_sweep(C.byref(spec.desc), _ptr(flat_params), C.byref(q), C.byref(bd), xp,
                                   _ptr(eps_base), _ptr(eps_score), _ptr(eps_av), seed, _ptr(r.avatars),
                                   _ptr(r.sampled_scores), _ptr(r.reconstructions), _ptr(r.betas), _ptr(r.coefs),
                                   _ptr(r.pvalues), _ptr(ws), ws.numel(), _stream()))
    r._keep = (xs, eps_base, eps_score, eps_av, ws)
    r._desc = q
    return r


def daa_sweep_layered(spec: PathSpec, flat_params, src, dst, n_samples, n_base, *, src_mod=0, dst_mod=1,
                      sample_latents=True, reg_method="hierarchical", seed=1037, val_begin=0, eps_base=None,
                      eps_score=None, eps_av=None, scores=None, materialize=True, workspace=None):
    """The DAA sweep for the architectures of the layered path (hidden decoder layers make the decoder non-affine and
    a per-sample output scale is not a constant: the shortcuts of the fused sweep -- mean noise row of the base passes,
    slopes by linearity from z -- do not apply).  Per validation TWO forward launches of the C-ABI forward:
      * the M base passes as one batch of M * N rows (workflow.py:388-398), then their mean over M;
      * the n_samples x n_scores perturbed forwards as one batch of N * C * J rows in the avatar tensor's own
        (subject, score, sample) order (workflow.py:406-419),
    each row selecting its mixture component like the row of the N-subject batch it stands for (batch desc
    owner_div / owner_mod); then the regression kernels read the materialised tensor (`daa_regression`).
    Same arguments / result as `daa_sweep`; the generator noise is keyed by (seed, global validation, stage)."""
    from . import engine
    if reg_method not in REG_METHODS:
        raise NotImplementedError("reg_method=%r is not on the B200 path (hierarchical, fixed)" % (reg_method,))
    if not materialize:
        raise NotImplementedError("the layered sweep regresses on the materialised avatar tensor")
    if spec.n_mods != 2:
        raise NotImplementedError("the layered sweep covers the reference's two-modality cohorts")
    _require_cuda(flat_params, "parameters")
    src, dst = _f32(src), _f32(dst)
    _require_cuda(src, "src")
    _require_cuda(dst, "dst")
    device = flat_params.device
    n_val, N, Cc = src.shape
    R, J, M, E = dst.shape[2], int(n_samples), int(n_base), spec.eps_width
    sname, dname = spec.mod_names[src_mod], spec.mod_names[dst_mod]
    if scores is not None and eps_score is not None:
        raise ValueError("scores and eps_score are exclusive")
    ws = workspace or Workspace()
    f = lambda dt, *s: torch.empty(*s, dtype=dt, device=device)
    r = DaaResult()
    r.avatars = f(torch.float32, n_val, N, Cc, J, R)
    r.sampled_scores = f(torch.float32, n_val, N, J, Cc)
    r.reconstructions = f(torch.float32, n_val, N, R)
    for v in range(n_val):
        key = (int(seed) * 1000003 + (val_begin + v)) & 0x7FFFFFFFFFFFFFFF
        # ---- M stochastic reconstructions (rows (pass, subject)) and their mean ----
        base = {sname: src[v].repeat(M, 1), dname: dst[v].repeat(M, 1)}
        eb = _f32(eps_base[v]).reshape(M * N, E) if eps_base is not None else None
        res = engine.forward(spec, flat_params, base, eps=eb, seed=3 * key, sample_latents=True, workspace=ws, owner=(1, N))
        loc_hat = res.rec_loc[src_mod].view(M, N, Cc).mean(0)
        lv = res.rec_logvar[src_mod] if spec.learn_output_sample_scale else flat_params[spec.layout.dec_lv[src_mod]:][:Cc].expand(M * N, Cc)
        scale_hat = (0.5 * lv).exp().view(M, N, Cc).mean(0)
        r.reconstructions[v] = res.rec_loc[dst_mod].view(M, N, R).mean(0)
        # ---- artificial scores (J, N, C) ----
        if scores is not None:
            sc = _f32(scores[v])
        else:
            if eps_score is not None:
                es = _f32(eps_score[v])
            else:
                es = f(torch.float32, J, N, Cc)
                _lib.check(_lib.lib().mopoe_philox_normal(3 * key + 1, _lib.STREAM_DAA_SCORE, 0, es.numel(), _ptr(es), _stream()))
            sc = loc_hat + scale_hat * es                                   # Normal(loc_hat, scale_hat).sample, workflow.py:401-405
        r.sampled_scores[v] = sc.permute(1, 0, 2)
        # ---- one perturbed forward per (subject, score, sample), rows in that order ----
        cdata = src[v][:, None, None, :].expand(N, Cc, J, Cc).clone()
        idx = torch.arange(Cc, device=device)
        cdata[:, idx, :, idx] = sc.permute(2, 1, 0)                          # [c, n, j] -> column c of rows (n, c, j)
        pert = {sname: cdata.view(N * Cc * J, Cc), dname: dst[v][:, None, :].expand(N, Cc * J, R).reshape(N * Cc * J, R)}
        ea = None
        if sample_latents and eps_av is not None:
            ea = _f32(eps_av[v]).permute(2, 1, 0, 3).reshape(N * Cc * J, E)   # (J, C, N, E) -> rows (n, c, j)
        res = engine.forward(spec, flat_params, pert, eps=ea, seed=3 * key + 2, sample_latents=sample_latents, workspace=ws,
                             owner=(Cc * J, N))
        r.avatars[v] = res.rec_loc[dst_mod].view(N, Cc, J, R)
    r.pvalues, r.coefs, r.betas = daa_regression(r.avatars, r.sampled_scores, r.reconstructions, reg_method=reg_method)
    r._keep, r._desc = (ws,), None
    return r


def check_status(spec: PathSpec, result):
    """Synchronise and raise MopoeError if the sweep that produced `result` flagged a device-side protocol error
    (a bounded tcgen05 / mbarrier wait timed out; the kernels then poison the tables with NaN rather than hang).
    Callers that persist results (workflow.daa_exp) call this before writing anything."""
    if result._desc is None:          # layered sweep: no bounded device-side waits to report on
        torch.cuda.synchronize()
        return result
    _lib.check(_lib.lib().mopoe_daa_status(C.byref(spec.desc), C.byref(result._desc), _ptr(result._keep[-1]), _stream()))
    return result


def phase_cycles(spec: PathSpec, result):
    """Per-phase cycle counters of the tcgen05 avatar kernel of `result`'s sweep (profiling aid)."""
    out = (C.c_int64 * 32)()
    torch.cuda.synchronize()
    _lib.check(_lib.lib().mopoe_daa_read_phases(C.byref(spec.desc), C.byref(result._desc), _ptr(result._keep[-1]), out))
    return list(out)


def daa_regression(avatars, sampled_scores, reconstructions=None, reg_method="hierarchical"):
    """stat_utils.make_regression for every (validation, score, roi) of a materialised avatar tensor
    (workflow.py:466-505).  avatars (n_val,N,C,J,R) fp32 CUDA, sampled_scores (n_val,N,J,C)."""
    if reg_method not in REG_METHODS:
        raise NotImplementedError("reg_method=%r is not on the B200 path (hierarchical, fixed)" % (reg_method,))
    avatars, sampled_scores = _f32(avatars), _f32(sampled_scores)
    _require_cuda(avatars, "avatars")
    n_val, N, Cc, J, R = avatars.shape
    device = avatars.device
    betas = torch.empty(n_val, Cc, N, R, dtype=torch.float64, device=device)
    coefs = torch.empty(n_val, Cc, R, dtype=torch.float64, device=device)
    pvalues = torch.empty(n_val, Cc, R, dtype=torch.float64, device=device)
    rec = _f32(reconstructions) if reconstructions is not None else None
    _lib.check(_lib.lib().mopoe_daa_regression(n_val, N, Cc, J, R, REG_METHODS[reg_method], _ptr(avatars),
                                               _ptr(sampled_scores), _ptr(rec), _ptr(betas), _ptr(coefs),
                                               _ptr(pvalues), _stream()))
    return pvalues, coefs, betas


def significant(pvalues, trust_level):
    """workflow.py:517-523.  pvalues (n_val, n_scores, n_rois) numpy/torch -> bool (n_scores, n_rois)."""
    p = pvalues.detach().cpu().numpy() if torch.is_tensor(pvalues) else np.asarray(pvalues)
    n_val, Cc, R = p.shape
    thr = 0.05 / R / Cc
    return (p < thr).sum(axis=0) >= trust_level * n_val


# ---- multi-GPU: validations are independent units --------------------------------
def shard_validations(n_validation, rank, world_size):
    """Contiguous block split of range(n_validation): -> (begin, end) of this rank."""
    base, rem = divmod(n_validation, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_units(n_validation, n_scores, rank, world_size):
    """contiguous block split of the n_validation * n_scores (validation, score) units (140 for the
    HBN sweep: 18 / 17 per rank on 8 GPUs, 97 % balance, where whole validations give 3 / 2 = 83 %).
    -> dict(unit_begin, unit_end: global units of this rank; val_begin, val_end: the validations it touches (their
    base passes are run by every rank that shares them); local_begin, local_end: the same units relative to
    val_begin, the values `daa_sweep(unit_begin=, unit_end=)` takes)."""
    ub, ue = shard_validations(n_validation * n_scores, rank, world_size)
    if ue == ub:
        return dict(unit_begin=ub, unit_end=ue, val_begin=0, val_end=0, local_begin=0, local_end=0)
    vb, ve = ub // n_scores, (ue + n_scores - 1) // n_scores
    return dict(unit_begin=ub, unit_end=ue, val_begin=vb, val_end=ve, local_begin=ub - vb * n_scores, local_end=ue - vb * n_scores)


def gather_tables(local, n_validation, group=None, out=None):
    """all_gather the per-validation fp64 tables of every rank into the full (n_validation, ...) table.
    `local` is this rank's (n_local, ...) block.  Equal shards (the usual case) are gathered straight into
    one tensor (`out` lets the caller reuse it: one NCCL kernel, nothing else); unequal shards are padded
    to the largest one."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    counts = [shard_validations(n_validation, r, world) for r in range(world)]
    sizes = [e - b for b, e in counts]
    if min(sizes) == max(sizes):
        shape = (n_validation,) + tuple(local.shape[1:])
        if out is None or tuple(out.shape) != shape or out.dtype != local.dtype or out.device != local.device:
            out = torch.empty(shape, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    return gather_rows(local, sizes, group)


def gather_rows(local, sizes, group=None):
    """all_gather of per-rank blocks of leading sizes `sizes` (known on every rank; zero allowed), padded to the
    largest: -> the concatenation in rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert len(sizes) == world and local.shape[0] == sizes[dist.get_rank(group)]
    n_max = max(max(sizes), 1)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: n] for r, n in enumerate(sizes)], dim=0)


def gather_tables_many(locals_, n_validation, group=None, outs=None):
    """gather_tables for several tables of equal shards in ONE coalesced NCCL launch (the sweep itself is
    under a millisecond, so each extra collective launch shows in the multi-GPU step time).  Falls back to
    one collective per table when the shards are unequal or coalescing is unavailable."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(locals_)
    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (shard_validations(n_validation, r, world) for r in range(world))]
    outs = list(outs) if outs is not None else [None] * len(locals_)
    # coalescing of all_gather_into_tensor is an NCCL feature (gloo silently mis-gathers under it)
    if min(sizes) != max(sizes) or not hasattr(dist, "_coalescing_manager") or dist.get_backend(group) != "nccl":
        return [gather_tables(t, n_validation, group, o) for t, o in zip(locals_, outs)]
    for i, t in enumerate(locals_):
        shape = (n_validation,) + tuple(t.shape[1:])
        if outs[i] is None or tuple(outs[i].shape) != shape or outs[i].dtype != t.dtype or outs[i].device != t.device:
            outs[i] = torch.empty(shape, dtype=t.dtype, device=t.device)
    try:
        with dist._coalescing_manager(group=group, device=locals_[0].device, async_ops=False):
            for t, o in zip(locals_, outs):
                dist.all_gather_into_tensor(o, t.contiguous(), group=group)
    except Exception:
        for t, o in zip(locals_, outs):
            dist.all_gather_into_tensor(o, t.contiguous(), group=group)
    return outs


class TableExchange:
    """The coefs / p-value tables of every rank's validations in EVERY rank's memory without a collective launch:
    each rank stores its (n_local, C, R) slices straight into all ranks' full tables through NVLink peer pointers
    (symmetric memory) and releases a sequence flag; a one-warp kernel acquires the flags of all sources
    (C-ABI mopoe_daa_exchange_tables; two small launches on the current stream, capturable in a CUDA graph).
    Replaces the NCCL all_gather of `gather_tables` on the sweep's critical path."""

    def __init__(self, n_val_total, n_scores, n_rois, device, group=None, root=None):
        """root=None: every rank ends up with the full tables (all-gather); root=r: only rank r does (gather, what
        daa_exp needs: 1/world of the bytes, and no rank but r ever waits for another)."""
        import torch.distributed as dist
        self.root = -1 if root is None else int(root)
        import torch.distributed._symmetric_memory as symm
        lib = _lib.lib()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > _lib.MAX_PEERS:
            raise _lib.MopoeError("table exchange supports up to %d ranks" % _lib.MAX_PEERS)
        self.shape = (int(n_val_total), int(n_scores), int(n_rois))
        self.elems_total = self.shape[0] * self.shape[1] * self.shape[2]
        nbytes = int(lib.mopoe_table_exchange_bytes(self.elems_total))
        try:
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        torch.cuda.synchronize()
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        dist.barrier(self.group)                      # every buffer is zeroed and mapped before anyone pushes
        torch.cuda.synchronize()
        self.calls = 0                                # launches so far == the device-side sequence number

    def exchange(self, coefs_local, pvalues_local, val_begin):
        """Push this rank's slices (n_local, C, R) that start at validation `val_begin`; returns the number of
        exchanges issued so far (its parity selects the slot that holds the result, see `tables`)."""
        c, p = coefs_local.contiguous(), pvalues_local.contiguous()
        assert c.dtype == torch.float64 and p.dtype == torch.float64 and tuple(c.shape[1:]) == self.shape[1:]
        per_val = self.shape[1] * self.shape[2]
        d = _lib.TableExchangeDesc(world=self.world, rank=self.rank, root=self.root, reserved=0, elems_local=c.shape[0] * per_val,
                                   elem_offset=int(val_begin) * per_val, elems_total=self.elems_total)
        for r, ptr in enumerate(self.peer_ptrs):
            d.peer_base[r] = ptr
        _lib.check(_lib.lib().mopoe_daa_exchange_tables(C.byref(d), _ptr(c), _ptr(p), _stream()))
        self._keep = (c, p)
        self.calls += 1
        return self.calls

    def exchange_units(self, coefs_local, pvalues_local, local_begin, local_end, unit_begin):
        """`exchange` for (validation, score) shard units (`shard_units`): rows [local_begin, local_end) of this rank's
        (n_val_local * C, R) tables are global units [unit_begin, unit_begin + local_end - local_begin)."""
        assert coefs_local.is_contiguous() and pvalues_local.is_contiguous() and coefs_local.dtype == torch.float64
        R = self.shape[2]
        c = coefs_local.view(-1, R)[local_begin:local_end]
        p = pvalues_local.view(-1, R)[local_begin:local_end]
        d = _lib.TableExchangeDesc(world=self.world, rank=self.rank, root=self.root, reserved=0, elems_local=c.shape[0] * R,
                                   elem_offset=int(unit_begin) * R, elems_total=self.elems_total)
        for r, ptr in enumerate(self.peer_ptrs):
            d.peer_base[r] = ptr
        _lib.check(_lib.lib().mopoe_daa_exchange_tables(C.byref(d), _ptr(c), _ptr(p), _stream()))
        self._keep = (coefs_local, pvalues_local)
        self.calls += 1
        return self.calls

    def tables(self, calls=None):
        """(coefs, pvalues) views, each (n_val_total, C, R) fp64, of the slot written by exchange number `calls`
        (default: the last one issued from Python; pass the replay count when the exchange runs inside a CUDA graph)."""
        calls = self.calls if calls is None else calls
        off = 256 + (calls & 1) * 2 * self.elems_total * 8
        flat = self.buf[off:off + 2 * self.elems_total * 8].view(torch.float64)
        return flat[:self.elems_total].view(self.shape), flat[self.elems_total:].view(self.shape)


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPU cores local to GPU `device_index` (NVML affinity mask) so that the pinned
    host buffers it allocates afterwards (the 1.9 GB avatar tensor of a sweep) land on that GPU's NUMA node:
    with one process per GPU and no binding, several ranks' device-to-host copies cross the socket
    interconnect and contend there.  Best effort: returns the core list, or None if NVML / the affinity call is
    unavailable (nothing changes then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None

