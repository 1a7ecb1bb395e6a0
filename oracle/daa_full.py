"""Full-size DAA sweep through the CPU oracle with the production (Philox) noise, one validation per worker
process.  TEST INFRASTRUCTURE (see oracle/__init__.py): used by the slow GPU parity test of the trained-model
HBN sweep (BASELINE.json configs[3]) and nowhere on the product path.

Each worker regenerates the noise of ITS validation with the numpy restatement of the device generator
(oracle/philox.py; rows / elements are addressed by the global validation index, exactly as the kernels do),
runs workflow.py:388-419 through daa_oracle.daa_generate and the closed-form hierarchical regression."""
import numpy as np
import torch

from . import daa_oracle, mopoe_oracle as mo, philox


def validation(args):
    """-> (v, avatars (N, C, J, R) float32, sampled (N, J, C), recon (N, R), pvalues (C, R), coefs (C, R))."""
    v, spec_kw, params_np, src_v, dst_v, seed, n_base, n_samples = args
    torch.set_num_threads(1)
    spec = mo.ModelSpec(**spec_kw)
    params = {k: torch.from_numpy(a) for k, a in params_np.items()}
    N, C = src_v.shape
    E, L, sd = spec.eps_width, spec.latent_dim, list(spec.style_dims)
    J, Mb = n_samples, n_base
    eb = torch.from_numpy(philox.philox_rows(seed, philox.STREAM_DAA_BASE, Mb * N, L, sd, row_start=v * Mb * N)).view(1, Mb, N, E)
    es = torch.from_numpy(philox.philox_normal(seed, philox.STREAM_DAA_SCORE, J * N * C, start=v * J * N * C)).view(1, J, N, C)
    ea = torch.from_numpy(philox.philox_rows(seed, philox.STREAM_DAA_AVATAR, J * C * N, L, sd, row_start=v * J * C * N)).view(1, J, C, N, E)
    av, sc, rc = daa_oracle.daa_generate(params, spec, torch.from_numpy(src_v)[None], torch.from_numpy(dst_v)[None], eb, es, ea)
    p, coef, _ = daa_oracle.hierarchical_regression(av, sc)
    return v, av[0], sc[0], rc[0], p[0], coef[0]


def sweep(spec_kw, params, src, dst, seed, n_base, n_samples, workers=None):
    """src (n_val, N, C), dst (n_val, N, R) numpy float32; params: dict of CPU tensors.  Yields the per-validation
    results as they finish (so the caller can compare and drop the 93 MB avatar block of each validation)."""
    import multiprocessing as mp
    import os
    params_np = {k: v.detach().cpu().numpy() for k, v in params.items()}
    jobs = [(v, spec_kw, params_np, np.ascontiguousarray(src[v]), np.ascontiguousarray(dst[v]), seed, n_base, n_samples)
            for v in range(src.shape[0])]
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2) - 1))
    ctx = mp.get_context("spawn")          # the parent holds a CUDA context: never fork it
    with ctx.Pool(workers) as pool:
        for res in pool.imap_unordered(validation, jobs):
            yield res
