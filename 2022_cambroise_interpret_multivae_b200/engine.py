"""Thin functional layer over the C-ABI: tensors in, tensors out, everything on the current CUDA
stream.  Used by the reference-interface mirrors (model.py, run_epochs.py, workflow.py)."""
import ctypes as C

import torch

from . import _lib
from .spec import PathSpec


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.MopoeError("%s must live on a CUDA device: the MoPoE B200 path has no CPU fallback" % what)


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


class Workspace:
    """Grow-only device scratch buffer."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


def init_params(spec: PathSpec, seed=0):
    """Random-init weights with torch's nn.Linear default distribution, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for
    weights and biases (networks.py:17,23-28,56 construct plain nn.Linear layers), and the decoder logvar filled
    with initial_out_logvar (networks.py:61-64), drawn from a private CPU generator in state-dict order.
    -> dict of state-dict tensors (CPU fp32).  Benchmarks and synthetic runs start from this."""
    import math
    g = torch.Generator().manual_seed(seed)
    slices = spec.param_slices()
    params = {}
    for name, (_, shape) in slices.items():
        if name.startswith("decoders.") and name.endswith("logvar"):
            params[name] = torch.full(shape, spec.initial_out_logvar, dtype=torch.float32)
            continue
        fan_in = shape[1] if name.endswith(".weight") else slices[name[:-4] + "weight"][1][1]
        bound = 1.0 / math.sqrt(fan_in)
        params[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(torch.float32)
    return params


def pack_params(spec: PathSpec, params, device):
    """dict of state-dict tensors -> flat fp32 parameter buffer in C-ABI layout."""
    flat = torch.zeros(spec.layout.total, dtype=torch.float32, device=device)
    for name, (off, shape) in spec.param_slices().items():
        n = 1
        for s in shape:
            n *= s
        flat[off:off + n].copy_(params[name].detach().reshape(-1).to(device=device, dtype=torch.float32))
    return flat


def unpack_params(spec: PathSpec, flat):
    out = {}
    for name, (off, shape) in spec.param_slices().items():
        n = 1
        for s in shape:
            n *= s
        out[name] = flat[off:off + n].view(shape)
    return out


class ForwardResult:
    """Raw outputs of one mopoe_forward / single-step mopoe_train_steps call."""

    def __init__(self, spec, n_rows, present_mask, device, want_rec=True):
        L, nsub = spec.latent_dim, len(spec.subsets())
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        self.present_mask = present_mask
        self.enc_heads = [f(n_rows, spec.head_cols(m)) if present_mask >> m & 1 else None for m in range(spec.n_mods)]
        self.subset_mu, self.subset_logvar = f(nsub, n_rows, L), f(nsub, n_rows, L)
        self.joint_mu, self.joint_logvar, self.z = f(n_rows, L), f(n_rows, L), f(n_rows, L)
        self.z_style = [f(n_rows, spec.style_dims[m]) if (present_mask >> m & 1) and spec.style_dims[m] > 0 else None
                        for m in range(spec.n_mods)]
        self.rec_loc = [f(n_rows, spec.dims[m]) if want_rec and (present_mask >> m & 1) else None
                        for m in range(spec.n_mods)]
        self.rec_logvar = [f(n_rows, spec.dims[m]) if want_rec and spec.learn_output_sample_scale and (present_mask >> m & 1) else None
                           for m in range(spec.n_mods)]          # per-sample output log-variance (networks.py:73-74)
        self.scalars = torch.zeros(_lib.N_SCALARS, dtype=torch.float32, device=device)

    def as_struct(self):
        o = _lib.ForwardOut()
        for m in range(len(self.enc_heads)):
            o.enc_heads[m] = _ptr(self.enc_heads[m]).value
            o.z_style[m] = _ptr(self.z_style[m]).value
            o.rec_loc[m] = _ptr(self.rec_loc[m]).value
            o.rec_logvar[m] = _ptr(self.rec_logvar[m]).value
        o.subset_mu, o.subset_logvar = _ptr(self.subset_mu), _ptr(self.subset_logvar)
        o.joint_mu, o.joint_logvar, o.z = _ptr(self.joint_mu), _ptr(self.joint_logvar), _ptr(self.z)
        o.scalars = _ptr(self.scalars)
        return o


def forward(spec: PathSpec, flat_params, batch, eps=None, seed=0, sample_latents=True, use_expert=None,
            with_nll=False, workspace=None, owner=None):
    """BaseMMVae.forward on the GPU.  batch: dict name -> (N, D_m) CUDA tensor (present only).
    owner=(div, P): the N rows are P-row reference batches laid out side by side (see PathSpec.batch_desc)."""
    _require_cuda(flat_params, "parameters")
    device = flat_params.device
    mask = spec.present_mask(batch.keys())
    xs = [(_f32(batch[n]) if n in batch else None) for n in spec.mod_names]
    for x in xs:
        if x is not None:
            _require_cuda(x, "input batch")
    n_rows = next(x for x in xs if x is not None).shape[0]
    bd = spec.batch_desc(n_rows, mask, owner=owner)
    if eps is not None:
        eps = _f32(eps)
        assert eps.shape == (n_rows, spec.eps_width), (eps.shape, (n_rows, spec.eps_width))
    ue = -1
    if use_expert is not None:
        keys = [k for k, _ in spec.subsets()]
        if use_expert not in keys:
            raise KeyError(use_expert)
        ue = keys.index(use_expert)
    res = ForwardResult(spec, n_rows, mask, device)
    lib = _lib.lib()
    nbytes = lib.mopoe_workspace_bytes(C.byref(spec.desc), n_rows)
    ws = (workspace or Workspace()).get(nbytes, device)
    xp = (C.c_void_p * _lib.MAX_MODS)(*[_ptr(x).value for x in xs] + [None] * (_lib.MAX_MODS - len(xs)))
    out = res.as_struct()
    _lib.check(lib.mopoe_forward(C.byref(spec.desc), _ptr(flat_params), C.byref(bd), xp, _ptr(eps), seed,
                                 int(bool(sample_latents)), ue, int(bool(with_nll)), C.byref(out), _ptr(ws),
                                 ws.numel(), _stream()))
    res._keep = (xs, eps, ws)
    return res


def make_batches(spec: PathSpec, batch_list, device):
    """batch_list: [(n_rows, present_mask, row_offset)] -> device array of mopoe_batch_desc."""
    arr = (_lib.BatchDesc * len(batch_list))(*[spec.batch_desc(n, p, o) for n, p, o in batch_list])
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    return raw


def train_steps(spec: PathSpec, flat_params, data, batches_dev, n_steps, max_rows, mode, row_index=None,
                eps=None, seed=0, adam_m=None, adam_v=None, adam_t=None, grads=None, lr=0.002, b1=0.9,
                b2=0.999, adam_eps=1e-8, forward_result=None, workspace=None):
    """n_steps x (basic_routine_epoch [+ backward [+ Adam]]) in one persistent launch.
    data: list of per-modality (rows, D_m) CUDA tensors (None where a modality has no data)."""
    _require_cuda(flat_params, "parameters")
    device = flat_params.device
    lib = _lib.lib()
    scalars = torch.zeros(n_steps, _lib.N_SCALARS, dtype=torch.float32, device=device)
    nbytes = lib.mopoe_workspace_bytes(C.byref(spec.desc), max_rows)
    ws = (workspace or Workspace()).get(nbytes, device)
    pad = [None] * (_lib.MAX_MODS - spec.n_mods)
    dp = (C.c_void_p * _lib.MAX_MODS)(*[_ptr(x).value for x in data] + pad)
    rp = None
    if row_index is not None:
        rp = (C.c_void_p * _lib.MAX_MODS)(*[_ptr(x).value for x in row_index] + pad)
    if eps is not None:
        assert eps.dtype == torch.float32 and eps.is_contiguous()
        assert eps.shape == (n_steps, spec.n_pass, max_rows, spec.eps_width), eps.shape
    out = forward_result.as_struct() if forward_result is not None else None
    _lib.check(lib.mopoe_train_steps(
        C.byref(spec.desc), _ptr(flat_params), _ptr(adam_m), _ptr(adam_v), _ptr(adam_t), _ptr(grads), dp, rp,
        _ptr(batches_dev), n_steps, max_rows, _ptr(eps), seed, mode, lr, b1, b2, adam_eps, _ptr(scalars),
        C.byref(out) if out is not None else None, _ptr(ws), ws.numel(), _stream()))
    return scalars


def philox_normal(seed, stream_id, n, device, start=0):
    out = torch.empty(n, dtype=torch.float32, device=device)
    _lib.check(_lib.lib().mopoe_philox_normal(seed, stream_id, start, n, _ptr(out), _stream()))
    return out
