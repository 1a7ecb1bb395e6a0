"""Representational similarity analysis on the GPU: device half of experiments/workflow.py:656-789
(rsa_exp) behind the reference's stat_utils names -- data2cmat / vec2cmat (stat_utils.py:25-33,46-53), cmat2triu
(:36-43) and fit_rsa (:81-95, scipy.stats.kendalltau of the two upper triangles).

The pairwise matrices and the P^2 = (n (n - 1) / 2)^2 entry-pair comparisons of Kendall's tau run in CUDA kernels
(csrc/mopoe_rsa.cu, exact integer counts); what is left for the host is the closed form that turns seven integers
into tau-b and its p-value, in the variant scipy's `method="auto"` selects."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .engine import Workspace, _ptr, _require_cuda, _stream


def data2cmat(data):
    """stat_utils.py:25-33: Euclidean distance matrix of the rows of `data` (n, d) -> (n, n) fp64 CUDA tensor
    (a (k, n, d) stack gives (k, n, n))."""
    data = torch.as_tensor(data)
    _require_cuda(data, "data")
    if data.ndim > 2:
        return torch.stack([data2cmat(x) for x in data])
    x = data.detach().to(torch.float32).contiguous()
    n, d = x.shape
    out = torch.empty(n, n, dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().mopoe_rsa_cmat(n, d, _ptr(x), 0, _ptr(out), _stream()))
    return out


def vec2cmat(vec, categorical=False, metric="euclidean"):
    """stat_utils.py:46-53: |v_a - v_b| (fp64) or, for a categorical characteristic, the 0 / 1 "differs" matrix.
    `vec`: a CUDA tensor, or anything numpy understands (categorical labels are coded by np.unique first)."""
    if metric != "euclidean":
        raise NotImplementedError("metric=%r is not on the B200 path (euclidean)" % (metric,))
    if not torch.is_tensor(vec):
        arr = np.asarray(vec)
        if categorical:
            arr = np.unique(arr, return_inverse=True)[1]
        vec = torch.as_tensor(arr.astype(np.float32))
        vec = vec.cuda()
    _require_cuda(vec, "vec")
    x = vec.detach().to(torch.float32).reshape(-1, 1).contiguous()
    n = x.shape[0]
    out = torch.empty(n, n, dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().mopoe_rsa_cmat(n, 1, _ptr(x), int(bool(categorical)), _ptr(out), _stream()))
    return out


def cmat2triu(arr):
    """stat_utils.py:36-43 (host-side convenience; the kernels take the matrices themselves)."""
    assert arr.ndim == 2 and arr.shape[0] == arr.shape[1]
    iu = torch.triu_indices(arr.shape[0], arr.shape[0], 1, device=arr.device)
    return arr[iu[0], iu[1]]


def kendall_counts(cmat, ref_cmats, workspace=None):
    """(n, n) fp64 x (n_ref, n, n) fp64 -> (n_ref, 7) int64 CUDA tensor (include/mopoe_b200.h: mopoe_rsa_kendall)."""
    _require_cuda(cmat, "cmat")
    _require_cuda(ref_cmats, "ref_cmats")
    cmat, ref_cmats = cmat.contiguous(), ref_cmats.contiguous()
    assert cmat.dtype == torch.float64 and ref_cmats.dtype == torch.float64
    n = cmat.shape[0]
    assert cmat.shape == (n, n) and ref_cmats.ndim == 3 and tuple(ref_cmats.shape[1:]) == (n, n)
    n_ref = ref_cmats.shape[0]
    lib = _lib.lib()
    nbytes = int(lib.mopoe_rsa_kendall_workspace_bytes(n, n_ref))
    if nbytes < 0:
        raise _lib.MopoeError("rsa_kendall: invalid sizes n=%d n_ref=%d" % (n, n_ref))
    ws = (workspace or Workspace()).get(nbytes, cmat.device)
    counts = torch.empty(n_ref, 7, dtype=torch.int64, device=cmat.device)
    _lib.check(lib.mopoe_rsa_kendall(n, n_ref, _ptr(cmat), _ptr(ref_cmats), _ptr(counts), _ptr(ws), ws.numel(), _stream()))
    return counts


def _exact_two_sided(size, c):
    """P(at most c discordant pairs or as extreme on the other side) for `size` untied observations: the number of
    permutations of `size` items with k inversions, built item by item (Kendall, Rank Correlation Methods)."""
    tot = size * (size - 1) // 2
    c = int(min(c, tot - c))
    if size <= 2:
        return 1.0
    if 4 * c == size * (size - 1):
        return 1.0
    if c <= 1 and size >= 171:
        return 0.0
    if size > 170:
        raise _lib.MopoeError("exact Kendall p-value for %d untied entries is not implemented" % size)
    ways = [1] + [0] * c                              # ways[k]: permutations of the first j items with k inversions
    for j in range(2, size + 1):
        nxt, run = [0] * (c + 1), 0
        for k in range(c + 1):
            run += ways[k]
            if k >= j:
                run -= ways[k - j]
            nxt[k] = run
        ways = nxt
    return min(1.0, max(0.0, 2.0 * sum(ways) / math.factorial(size)))


def kendall_from_counts(counts, size):
    """Seven integers per reference -> (tau, pvalue) arrays as scipy.stats.kendalltau(x, y) returns them
    (variant "b", two-sided, method "auto": exact without ties when size <= 33 or at most one discordant /
    concordant pair, else the normal approximation with the tie-corrected variance)."""
    c = np.asarray(counts.cpu() if torch.is_tensor(counts) else counts, dtype=np.int64).reshape(-1, 7)
    taus, pvals = np.empty(len(c)), np.empty(len(c))
    tot = size * (size - 1) // 2
    for r, row in enumerate(c):
        cmd2, sx, x0, x1, sy, y0, y1 = (int(v) for v in row)
        con_minus_dis, xtie, ytie = cmd2 // 2, sx // 2, sy // 2
        if xtie == tot or ytie == tot:
            taus[r], pvals[r] = np.nan, np.nan
            continue
        tau = con_minus_dis / np.sqrt(tot - xtie) / np.sqrt(tot - ytie)
        taus[r] = min(1.0, max(-1.0, tau))
        if xtie == 0 and ytie == 0:
            dis = (tot - con_minus_dis) // 2
            if size <= 33 or min(dis, tot - dis) <= 1:
                pvals[r] = _exact_two_sided(size, dis)
                continue
        m = size * (size - 1.0)
        var = (m * (2 * size + 5) - x1 - y1) / 18 + (2 * xtie * ytie) / m + x0 * y0 / (9 * m * (size - 2))
        pvals[r] = math.erfc(abs(con_minus_dis / math.sqrt(var)) / math.sqrt(2.0))
    return taus, pvals


def fit_rsa(cmat, ref_cmat, idxs=None, workspace=None):
    """stat_utils.py:81-95 for 2-D matrices: -> (tau, pval).  `ref_cmat` may be a stack (n_ref, n, n): -> arrays."""
    if cmat.ndim > 2:
        raise NotImplementedError("fit_rsa on a stack of matrices (the reference's debugging branch) is not on the B200 path")
    refs = ref_cmat if ref_cmat.ndim == 3 else ref_cmat[None]
    n = cmat.shape[0]
    taus, pvals = kendall_from_counts(kendall_counts(cmat, refs.to(torch.float64), workspace), n * (n - 1) // 2)
    return (taus, pvals) if ref_cmat.ndim == 3 else (float(taus[0]), float(pvals[0]))
