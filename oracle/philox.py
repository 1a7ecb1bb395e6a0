"""numpy restatement of the counter-based N(0,1) generator used by the CUDA path in production
mode (csrc/mopoe_rng.cuh).  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws its noise from torch's global generator (BaseMMVae.py:37-40,
workflow.py:401-405); that stream cannot be reproduced across devices, so parity is defined on
INJECTED noise.  Production mode replaces the injected tensor by
    eps[i] = philox_normal(seed, stream, i)         i = flat index into the injected layout
so "production mode" == "injected mode fed with this tensor", for any sharding of the sweep.
The two latent-noise tensors of the DAA sweep (eps_base, eps_av) are addressed by ROW instead of by
flat element (philox_rows below): a row is a whole number of Philox blocks, so the CUDA thread that
owns an avatar row draws whole blocks.

Algorithm: Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
counter = (i>>2 lo, i>>2 hi, stream lo, stream hi); the four 32-bit outputs become two
Box-Muller pairs: u = ((x >> 8) + 0.5) * 2^-24,  r = sqrt(-2 ln u_a),  (r cos 2 pi u_b, r sin 2 pi u_b).
Element i takes lane i & 3 (lanes 0,1 = cos,sin of pair (x0,x1); lanes 2,3 of pair (x2,x3)).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments uint32 arrays (broadcastable); returns four uint32 arrays."""
    c0 = np.asarray(c0, np.uint32); c1 = np.asarray(c1, np.uint32)
    c2 = np.asarray(c2, np.uint32); c3 = np.asarray(c3, np.uint32)
    k0 = np.asarray(k0, np.uint32); k1 = np.asarray(k1, np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return c0, c1, c2, c3


def philox_normal(seed, stream, n, start=0):
    """float32 array of n standard normals: elements start .. start+n-1 of (seed, stream)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    blk = idx >> np.uint64(2)
    lane = (idx & np.uint64(3)).astype(np.int64)
    seed = np.uint64(seed); stream = np.uint64(stream)
    x = philox4x32_10((blk & _MASK).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
                      np.uint32(stream & _MASK), np.uint32(stream >> np.uint64(32)),
                      np.uint32(seed & _MASK), np.uint32(seed >> np.uint64(32)))
    x = np.stack(x, axis=-1)                                   # (n, 4)
    u = ((x >> np.uint32(8)).astype(np.float64) + 0.5) * (1.0 / 16777216.0)
    pair = lane >> 1
    rows = np.arange(n)
    ua = u[rows, 2 * pair]
    ub = u[rows, 2 * pair + 1]
    # float32 arithmetic like the device (logf / sqrtf / sincospif), up to library ulps
    ua32 = ua.astype(np.float32); ub32 = ub.astype(np.float32)
    r = np.sqrt(np.float32(-2.0) * np.log(ua32)).astype(np.float32)
    ang = (np.float64(2.0) * ub32.astype(np.float64)) * np.pi
    trig = np.where((lane & 1) == 0, np.cos(ang), np.sin(ang)).astype(np.float32)
    return (r * trig).astype(np.float32)


def philox_rows(seed, stream, n_rows, latent_dim, style_dims, row_start=0):
    """(n_rows, E) float32 latent-noise rows of the DAA streams, E = latent_dim + sum(style_dims).
    Row r is drawn as EP/4 whole Philox blocks (counter = r * EP/4 + b) laid out
    [content | style_0 | style_1 ...] with every section padded to a multiple of 4 normals; the padding
    draws are discarded (csrc/mopoe_daa.cu: fill_noise_row)."""
    pad = lambda n: (n + 3) & ~3
    widths = [latent_dim] + list(style_dims)
    ep = sum(pad(w) for w in widths)
    flat = philox_normal(seed, stream, n_rows * ep, start=row_start * ep).reshape(n_rows, ep)
    cols, off = [], 0
    for w in widths:
        cols.append(flat[:, off:off + w])
        off += pad(w)
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


# stream ids shared with csrc/mopoe_rng.cuh
STREAM_DAA_BASE = 1      # eps_base  [n_val, M, N, E]                     rows: philox_rows
STREAM_DAA_SCORE = 2     # eps_score [n_val, n_samples, N, n_scores]
STREAM_DAA_AVATAR = 3    # eps_av    [n_val, n_samples, n_scores, N, E]   rows: philox_rows
STREAM_TRAIN = 4         # eps       [n_steps, n_pass, N, E]
STREAM_FORWARD = 5       # eps       [N, E]
