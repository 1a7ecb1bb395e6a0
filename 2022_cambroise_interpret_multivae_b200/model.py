"""Drop-in for the reference model class `multimodal_cohort.networks.VAE.VAE`
(experiments/multimodal_cohort/networks/VAE.py:6-8 -> experiments/utils/BaseMMVae.py).

Same constructor `VAE(flags, modalities, subsets)`, same state-dict keys and shapes
(`encoders.<m>.shared_encoder.0.weight` ... `decoders.<m>.out_mu.bias`), same
`forward(input_batch, sample_latents=True, use_expert=None) -> results` structure
(BaseMMVae.py:137-165) -- but the arithmetic runs in the sm_100a kernels behind
include/mopoe_b200.h.  Parameters are nn.Parameters whose storage is ONE flat fp32 buffer in the
C-ABI layout, so torch optimisers / checkpoints keep working and the fused Adam kernel can update
them in place.  The tensors in `results` are detached: training goes through
`run_epochs.basic_routine_epoch`, whose `total_loss` carries the kernel-computed gradients.
There is no CPU fallback: calling the model with CPU tensors raises.
"""
import os

import torch
import torch.nn as nn

from . import _lib, engine
from .spec import PathSpec


class Encoder(nn.Module):
    """Parameter container mirroring networks.py:9-28 (the math lives in the CUDA kernels)."""

    def __init__(self, spec: PathSpec, m):
        super().__init__()
        D, S, L = spec.dims[m], spec.style_dims[m], spec.latent_dim
        self.shared_encoder = nn.Sequential()
        width = D
        for _ in range(spec.n_hidden_enc):                       # networks.py:16-20
            self.shared_encoder.append(nn.Linear(width, _lib.HIDDEN))
            self.shared_encoder.append(nn.ReLU())
            self.shared_encoder.append(nn.Dropout(0.0))
            width = _lib.HIDDEN
        self.style_dim = S
        self.class_mu = nn.Linear(width, L)
        self.class_logvar = nn.Linear(width, L)
        if S > 0:
            self.style_mu = nn.Linear(width, S)
            self.style_logvar = nn.Linear(width, S)


class Decoder(nn.Module):
    """Parameter container mirroring networks.py:44-64."""

    def __init__(self, spec: PathSpec, m):
        super().__init__()
        D, S, L = spec.dims[m], spec.style_dims[m], spec.latent_dim
        self.style_dim = S
        self.shared_decoder = nn.Sequential()
        width = S + L
        for _ in range(spec.n_hidden_dec):                       # networks.py:51-55
            self.shared_decoder.append(nn.Linear(width, _lib.HIDDEN))
            self.shared_decoder.append(nn.ReLU())
            self.shared_decoder.append(nn.Dropout(0.0))
            width = _lib.HIDDEN
        self.out_mu = nn.Linear(width, D)
        if spec.learn_output_sample_scale:                       # networks.py:58-59
            self.logvar = nn.Linear(width, D)
        else:
            self.logvar = nn.Parameter(torch.full((1, D), spec.initial_out_logvar), requires_grad=spec.learn_output_scale)


class VAE(nn.Module):
    def __init__(self, flags, modalities, subsets=None):
        super().__init__()
        self.flags = flags
        self.modalities = modalities
        self.num_modalities = len(modalities)
        self.spec = PathSpec.from_flags(flags, mod_names=list(modalities.keys()))
        self.subsets = subsets if subsets is not None else {k: None for k, _ in self.spec.subsets()}
        encoders, decoders = nn.ModuleDict(), nn.ModuleDict()
        for m, key in enumerate(modalities.keys()):   # same construction order as BaseMMVae.py:28-31
            encoders[key] = Encoder(self.spec, m)
            decoders[key] = Decoder(self.spec, m)
        self.encoders, self.decoders = encoders, decoders
        lhood = {"normal": torch.distributions.Normal, "laplace": torch.distributions.Laplace}[self.spec.likelihood]
        self.lhoods = {k: lhood for k in modalities}             # modalities/modality.py:18-30
        self._flat = None
        self._noise = None
        self._ws = engine.Workspace()
        self.philox_seed = None      # set to an int to use the in-kernel generator instead of torch.randn
        self.rng_mode = "reference"   # "reference": torch.randn in the reference's call order; "block": one (n_pass, N, E) draw
        self._philox_calls = 0       # every forward / ELBO call draws fresh noise: the call count is folded into the seed

    # ---- flat parameter storage -------------------------------------------------------------
    def flat_parameters(self):
        """The flat fp32 parameter buffer (C-ABI layout) that every nn.Parameter is a view of."""
        named = dict(self.named_parameters())
        slices = self.spec.param_slices()
        dev = next(iter(named.values())).device
        if dev.type != "cuda":
            raise _lib.MopoeError("the model must be on a CUDA device: the MoPoE B200 path has no CPU fallback")
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            base = self._flat.data_ptr()
            ok = all(named[k].data_ptr() == base + 4 * off and named[k].dtype == torch.float32 for k, (off, _) in slices.items())
        if not ok:
            flat = torch.zeros(self.spec.layout.total, dtype=torch.float32, device=dev)
            for k, (off, shape) in slices.items():
                n = named[k].numel()
                flat[off:off + n].copy_(named[k].detach().reshape(-1).float())
                named[k].data = flat[off:off + n].view(shape)
            self._flat = flat
        return self._flat

    def inject_noise(self, eps):
        """Next forward/step consumes this (N, E) [or (n_pass, N, E)] tensor instead of fresh draws."""
        self._noise = eps

    def _next_seed(self):
        """Seed of the next in-kernel noise draw.  The generator is counter based (element index = step, pass,
        row, column), so a constant seed would hand every call the same noise: each call gets its own key,
        derived from (philox_seed, number of calls so far) -- reproducible for a given philox_seed."""
        if self.philox_seed is None:
            return 0
        seed = (int(self.philox_seed) + 0x9E3779B97F4A7C15 * self._philox_calls) & 0xFFFFFFFFFFFFFFFF
        self._philox_calls += 1
        return seed

    def _draw(self, n_pass, n_rows, device, batch_keys=None):
        """Noise of one forward / ELBO call as an (n_pass, N, E) tensor (E = content | style_0 | style_1 ...).
        Default (`rng_mode == "reference"`): torch.randn calls in the shapes and ORDER in which the reference consumes
        the global generator, so a same-seed run on the same device reproduces a reference run without injection:
        BaseMMVae.forward draws the joint content noise (N, L) first (BaseMMVae.py:143-144), then one (N, S_m) tensor
        per present modality with a style block, in `modalities` order (:155-159); in poe mode basic_routine_epoch
        then runs one unimodal forward per key of the batch dict (run_epochs.py:108-118), each drawing (N, L), (N, S_m)."""
        if self._noise is not None:
            eps, self._noise = self._noise, None
            return eps.to(device=device, dtype=torch.float32).reshape(n_pass, n_rows, self.spec.eps_width).contiguous()
        if self.philox_seed is not None:
            return None
        spec = self.spec
        if self.rng_mode != "reference" or batch_keys is None:
            return torch.randn(n_pass, n_rows, spec.eps_width, device=device)   # one draw from the global torch generator
        L = spec.latent_dim
        eps = torch.zeros(n_pass, n_rows, spec.eps_width, device=device)
        present = [m for m, n in enumerate(spec.mod_names) if n in batch_keys]

        def one(p, mods):
            eps[p, :, :L] = torch.randn(n_rows, L, device=device)
            for m in mods:
                S = spec.style_dims[m]
                if S > 0:
                    o = spec.style_offset(m)
                    eps[p, :, o:o + S] = torch.randn(n_rows, S, device=device)
        one(0, present)
        if n_pass > 1:                                   # poe: unimodal forwards in the order of the batch dict's keys
            for name in batch_keys:
                if name in spec.mod_names:
                    m = spec.mod_names.index(name)
                    one(1 + m, [m])
        return eps

    # ---- reference API ----------------------------------------------------------------------
    def reparameterize(self, mu, logvar):
        std = logvar.mul(0.5).exp()
        return torch.randn_like(std).mul(std).add(mu)            # BaseMMVae.py:37-40

    def _results(self, res, batch, scale_only_present=True):
        spec, L = self.spec, self.spec.latent_dim
        keys = [k for k, _ in spec.subsets()]
        avail, mix = spec.mixture_subsets(res.present_mask)
        enc = {}
        for m, name in enumerate(spec.mod_names):
            if res.enc_heads[m] is None:
                enc[name + "_style"] = [None, None]
                enc[name] = [None, None]
                continue
            h, S = res.enc_heads[m], spec.style_dims[m]
            enc[name] = [h[:, :L], h[:, L:2 * L]]
            enc[name + "_style"] = [h[:, 2 * L:2 * L + S], h[:, 2 * L + S:]] if S > 0 else [None, None]
        idx = torch.tensor(mix, device=res.subset_mu.device)
        mus, logvars = res.subset_mu.index_select(0, idx), res.subset_logvar.index_select(0, idx)
        sc = res.scalars
        ind, dyn_prior = sc[_lib.S_KLD_SUBSET:_lib.S_KLD_SUBSET + len(keys)].index_select(0, idx), None
        if spec.method == "jsd":                   # the prior is one more mixture component (BaseMMVae.py:217-223)
            zeros = torch.zeros_like(mus[:1])
            mus, logvars = torch.cat((mus, zeros)), torch.cat((logvars, zeros))
            ind = sc[_lib.S_JSD_DIV:_lib.S_JSD_DIV + len(mix) + 1]
            T = 1.0 / (logvars.exp() + 1e-8)       # results["dyn_prior"]: alpha_poe of the components (mm_div.py:23-35);
            pd_var = 1.0 / (T.mean(0))             # host-side view of what the kernel's divergence is measured against
            dyn_prior = [pd_var * (mus * T).mean(0), pd_var.log()]
        K = mus.shape[0]
        latents = {"modalities": enc, "mus": mus, "logvars": logvars,
                   "weights": (1 / float(K)) * torch.ones(K, device=idx.device),
                   "joint": [res.joint_mu, res.joint_logvar],
                   "subsets": {keys[s]: [res.subset_mu[s], res.subset_logvar[s]] for s in avail}}
        results = {"latents": latents, "group_distr": latents["joint"], "joint_divergence": sc[_lib.S_JOINT_DIV],
                   "individual_divs": ind, "dyn_prior": dyn_prior}
        rec = {}
        for m, name in enumerate(spec.mod_names):
            if res.rec_loc[m] is not None:
                lv = res.rec_logvar[m] if spec.learn_output_sample_scale else self.decoders[name].logvar.detach()
                rec[name] = self.lhoods[name](res.rec_loc[m], (lv * 0.5).exp())
        results["rec"] = rec
        results["class_embeddings"] = res.z
        return results

    def _forward_raw(self, input_batch, sample_latents=True, use_expert=None, with_nll=False):
        flat = self.flat_parameters()
        n_rows = len(next(iter(input_batch.values())))
        eps = self._draw(1, n_rows, flat.device, list(input_batch.keys())) if sample_latents else None
        seed = self._next_seed() if (sample_latents and eps is None) else 0
        return engine.forward(self.spec, flat, input_batch, eps=None if eps is None else eps[0], seed=seed,
                              sample_latents=sample_latents, use_expert=use_expert, with_nll=with_nll,
                              workspace=self._ws)

    @torch.no_grad()
    def forward(self, input_batch, sample_latents=True, use_expert=None):
        return self._results(self._forward_raw(input_batch, sample_latents, use_expert), input_batch)

    @torch.no_grad()
    def inference(self, input_batch, num_samples=None, sample=True, use_expert=None):
        return self.forward(input_batch, sample_latents=sample, use_expert=use_expert)["latents"]

    @torch.no_grad()
    def encode(self, input_batch):
        return self.forward(input_batch, sample_latents=False)["latents"]["modalities"]

    def save_networks(self):                                   # BaseMMVae.py:315-322
        for key in self.modalities:
            torch.save(self.encoders[key].state_dict(), os.path.join(self.flags.dir_checkpoints, "enc_" + key))
            torch.save(self.decoders[key].state_dict(), os.path.join(self.flags.dir_checkpoints, "dec_" + key))


class _ElboFunction(torch.autograd.Function):
    """total_loss with kernel-computed gradients: backward hands each nn.Parameter its slice of the
    flat gradient buffer (None for the parameters of absent modalities, as autograd would)."""

    @staticmethod
    def forward(ctx, loss, grads, *params):
        ctx.g = grads
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, gout):
        return (None, None) + tuple((None if g is None else g * gout) for g in ctx.g)


def elbo_step(model: VAE, input_batch, need_grad=True):
    """One basic_routine_epoch on the GPU (single cooperative launch): returns
    (scalars row, ForwardResult, total_loss tensor wired to the parameters)."""
    spec = model.spec
    flat = model.flat_parameters()
    dev = flat.device
    mask = spec.present_mask(input_batch.keys())
    data = [engine._f32(input_batch[n]) if n in input_batch else None for n in spec.mod_names]
    for x in data:
        if x is not None:
            engine._require_cuda(x, "input batch")
    n_rows = next(x for x in data if x is not None).shape[0]
    eps = model._draw(spec.n_pass, n_rows, dev, list(input_batch.keys()))
    res = engine.ForwardResult(spec, n_rows, mask, dev)
    bdev = engine.make_batches(spec, [(n_rows, mask, 0)], dev)
    grads = torch.zeros_like(flat) if need_grad else None
    sc = engine.train_steps(spec, flat, data, bdev, 1, n_rows, 1 if need_grad else 0,
                            eps=None if eps is None else eps[None].contiguous(),
                            seed=model._next_seed() if eps is None else 0,
                            grads=grads, forward_result=res, workspace=model._ws)[0]
    res.scalars = sc
    loss = sc[_lib.S_TOTAL_LOSS]
    if need_grad:
        named = dict(model.named_parameters())
        names = [k for k in spec.param_slices() if named[k].requires_grad]
        gviews = engine.unpack_params(spec, grads)
        glist = [gviews[k] if (mask >> spec.modality_of_param(k) & 1) else None for k in names]
        loss = _ElboFunction.apply(loss, glist, *[named[k] for k in names])
    return sc, res, loss
