"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# parity tolerance of the hot path (BASELINE.json north_star): 1e-4 relative in fp32
RTOL = 1e-4


def load_golden():
    with open(os.path.join(GOLDEN, "reference_elbo_forward.json")) as f:
        return json.load(f)


def digest(t):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    f = t.detach().reshape(-1).double().cpu()
    return [float(f.sum()), float(f.abs().sum())] + [float(v) for v in f[:6]]


def assert_digest_close(got, want, rtol=RTOL, what=""):
    """digest = [sum, abs_sum, first six values]; everything is compared relative to the
    tensor's own scale (abs_sum / numel is not stored, so use abs_sum as the scale of the sums
    and max(|first values|) for the leading entries)."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    scale_sum = max(abs(want[1]), 1e-30)
    assert abs(got[0] - want[0]) <= rtol * scale_sum, (what, "sum", got[0], want[0])
    assert abs(got[1] - want[1]) <= rtol * scale_sum, (what, "abs_sum", got[1], want[1])
    lead = max(np.abs(want[2:]).max(), 1e-30)
    assert np.all(np.abs(got[2:] - want[2:]) <= rtol * lead + 1e-7), (what, got[2:], want[2:])


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
