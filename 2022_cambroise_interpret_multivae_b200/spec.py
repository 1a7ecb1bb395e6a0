"""Host-side description of the hot path: the fields of the reference `flags` namespace
(experiments/workflow.py:98-149) that the MoPoE path reads, the subset table
(experiments/utils/BaseExperiment.py:58-79) and the bit-faithful selection boundaries
(experiments/utils/utils.py:63-85)."""
import ctypes as C
from itertools import chain, combinations

import torch

from . import _lib


def selection_bounds(n_rows, n_comp):
    """Row boundaries of utils.mixture_component_selection for the uniform weights the model uses,
    evaluated with the reference's own fp32 torch expression (BaseMMVae.py:225,99; utils.py:71-82):
    N=256,K=3 -> 85/85/86, N=50,K=3 -> 16/16/18."""
    w = (1 / float(n_comp)) * torch.ones(n_comp)
    w = w / w.sum()
    bounds = [0]
    for k in range(n_comp):
        if k == n_comp - 1:
            bounds.append(n_rows)
        else:
            bounds.append(bounds[-1] + int(torch.floor(n_rows * w[k])))
    bounds[-1] = n_rows
    return bounds


class PathSpec:
    def __init__(self, dims, style_dims, latent_dim=20, method="joint_elbo", mod_names=None,
                 learn_output_scale=True, beta=1.0, beta_style=1.0, beta_content=1.0,
                 num_hidden_layer_encoder=1, num_hidden_layer_decoder=0, likelihood="normal",
                 learn_output_sample_scale=False, initial_out_logvar=-3.0, dropout_rate=0.0):
        self.dims = [int(d) for d in dims]
        self.style_dims = [int(s) for s in style_dims]
        self.latent_dim = int(latent_dim)
        self.method = method
        names = list(mod_names) if mod_names is not None else ["clinical", "rois", "modc", "modd"]
        self.mod_names = names[: len(self.dims)]
        self.learn_output_scale = bool(learn_output_scale)
        self.beta, self.beta_style, self.beta_content = float(beta), float(beta_style), float(beta_content)
        self.n_hidden_enc = int(num_hidden_layer_encoder)
        self.n_hidden_dec = int(num_hidden_layer_decoder)
        self.likelihood = likelihood
        self.learn_output_sample_scale = bool(learn_output_sample_scale)
        self.initial_out_logvar = float(initial_out_logvar)
        if method not in _lib.METHODS:
            raise NotImplementedError("method=%r is not on the B200 path (poe, moe, joint_elbo, jsd)" % (method,))
        if likelihood not in _lib.LIKELIHOODS:
            raise NotImplementedError("likelihood=%r is not on the B200 path (normal, laplace)" % (likelihood,))
        if dropout_rate:
            raise NotImplementedError("dropout_rate != 0 is not on the B200 path")
        if len(self.style_dims) != len(self.dims) or len(self.mod_names) != len(self.dims):
            raise ValueError("dims / style_dims / mod_names length mismatch")
        self._desc = self._make_desc()
        self._layout = _lib.ParamLayout()
        # validation + layout arithmetic are pure host code: this works without a GPU
        _lib.check(_lib.lib().mopoe_param_layout_of(C.byref(self._desc), C.byref(self._layout)))

    @classmethod
    def from_flags(cls, flags, mod_names=None):
        """flags: the SimpleNamespace train_exp builds (workflow.py:98-149)."""
        method = ("poe" if getattr(flags, "modality_poe", False) else
                  "moe" if getattr(flags, "modality_moe", False) else
                  "jsd" if getattr(flags, "modality_jsd", False) else
                  "joint_elbo" if getattr(flags, "joint_elbo", False) else getattr(flags, "method", None))
        style = list(flags.style_dim) if not isinstance(flags.style_dim, int) else [flags.style_dim] * len(flags.input_dim)
        if len(style) != len(flags.input_dim):                       # experiment.py:133-136
            style = [style[0]] * len(flags.input_dim)
        if not flags.factorized_representation:
            style = [0] * len(flags.input_dim)
        return cls(flags.input_dim, style, flags.class_dim, method, mod_names,
                   learn_output_scale=flags.learn_output_scale, beta=flags.beta,
                   beta_style=flags.beta_style, beta_content=flags.beta_content,
                   num_hidden_layer_encoder=flags.num_hidden_layer_encoder,
                   num_hidden_layer_decoder=flags.num_hidden_layer_decoder, likelihood=flags.likelihood,
                   learn_output_sample_scale=getattr(flags, "learn_output_sample_scale", False),
                   initial_out_logvar=flags.initial_out_logvar,
                   dropout_rate=getattr(flags, "dropout_rate", 0.0))

    # ---- derived -------------------------------------------------------------------------
    @property
    def n_mods(self):
        return len(self.dims)

    @property
    def eps_width(self):
        return self.latent_dim + sum(self.style_dims)

    @property
    def n_pass(self):
        return 1 + self.n_mods if self.method == "poe" else 1

    def style_offset(self, m):
        return self.latent_dim + sum(self.style_dims[:m])

    def head_cols(self, m):
        return 2 * self.latent_dim + 2 * self.style_dims[m]

    def subsets(self):
        """[(key, [member modality indices in fusion order])] in set_subsets order, '' excluded."""
        out = []
        idx = list(range(self.n_mods))
        for combo in chain.from_iterable(combinations(idx, n) for n in range(1, len(idx) + 1)):
            names = sorted(self.mod_names[i] for i in combo)
            out.append(("_".join(names), [self.mod_names.index(n) for n in names]))
        return out

    def present_mask(self, keys):
        mask = 0
        for m, n in enumerate(self.mod_names):
            if n in keys:
                mask |= 1 << m
        if mask == 0:
            raise ValueError("input batch holds none of the modalities %s" % self.mod_names)
        return mask

    def mixture_subsets(self, present_mask):
        """indices (into subsets()) of the subsets that enter the mixture for this batch
        (fusion_condition_*, BaseMMVae.py:125-134) and of all available ones."""
        avail, mix = [], []
        n_present = bin(present_mask).count("1")
        for s, (_, members) in enumerate(self.subsets()):
            if any(not (present_mask >> m & 1) for m in members):
                continue
            avail.append(s)
            if self.method in ("moe", "jsd"):
                cond = len(members) == 1
            elif self.method == "poe":
                cond = len(members) == n_present
            else:
                cond = True
            if cond:
                mix.append(s)
        return avail, mix

    def _make_desc(self):
        d = _lib.ModelDesc()
        d.n_mods = self.n_mods
        order = sorted(range(self.n_mods), key=lambda m: self.mod_names[m])
        for m in range(self.n_mods):
            d.dims[m] = self.dims[m]
            d.style_dims[m] = self.style_dims[m]
            d.name_rank[m] = order.index(m)
        d.latent_dim = self.latent_dim
        d.hidden = _lib.HIDDEN
        d.n_hidden_enc = self.n_hidden_enc
        d.n_hidden_dec = self.n_hidden_dec
        d.method = _lib.METHODS[self.method]
        d.likelihood = _lib.LIKELIHOODS[self.likelihood]
        d.scale_mode = 1 if self.learn_output_sample_scale else 0
        d.learn_output_scale = int(self.learn_output_scale)
        d.beta, d.beta_style, d.beta_content = self.beta, self.beta_style, self.beta_content
        return d

    @property
    def desc(self):
        return self._desc

    @property
    def layout(self):
        return self._layout

    def batch_desc(self, n_rows, present_mask, row_offset=0, owner=None):
        """owner=(div, P): the rows are P-row reference batches side by side, row n selects its mixture component like
        row (n // div) % P of a P-row batch (include/mopoe_b200.h: mopoe_batch_desc.owner_div / owner_mod)."""
        b = _lib.BatchDesc()
        b.n_rows = int(n_rows)
        b.present_mask = int(present_mask)
        _, mix = self.mixture_subsets(present_mask)
        b.n_mix = len(mix) + (1 if self.method == "jsd" else 0)     # jsd: + the prior component (BaseMMVae.py:217-223)
        n_sel = int(owner[1]) if owner else int(n_rows)
        for i, v in enumerate(selection_bounds(n_sel, b.n_mix)):
            b.joint_bounds[i] = v
        for k in range(1, self.n_mods + 1):
            for i, v in enumerate(selection_bounds(n_sel, k)):
                b.moe_bounds[k][i] = v
        b.row_offset = int(row_offset)
        b.owner_div, b.owner_mod = (int(owner[0]), int(owner[1])) if owner else (0, 0)
        return b

    @property
    def layered(self):
        """True for the architectures that run on the layered path (csrc/mopoe_generic.cuh) instead of the fused
        kernels: hidden-layer counts other than (1, 0), per-sample output scale, non-normal likelihood."""
        return (self.n_hidden_enc != 1 or self.n_hidden_dec != 0 or self.learn_output_sample_scale
                or self.likelihood != "normal")

    def param_slices(self):
        """state-dict name -> (offset, shape) inside the flat parameter buffer, in the reference's own state-dict order
        (networks.py:9-28,44-64: Sequential index 3 l for the l-th hidden Linear, heads, out_mu, logvar)."""
        lay, L, H, out = self._layout, self.latent_dim, _lib.HIDDEN, {}
        He, Hd = self.n_hidden_enc, self.n_hidden_dec
        for m, name in enumerate(self.mod_names):
            D, S = self.dims[m], self.style_dims[m]
            e = "encoders.%s." % name
            for l in range(He):
                w, b = (lay.enc_w1[m], lay.enc_b1[m]) if l == 0 else (lay.enc_wx[m][l - 1], lay.enc_bx[m][l - 1])
                out[e + "shared_encoder.%d.weight" % (3 * l)] = (w, (H, D if l == 0 else H))
                out[e + "shared_encoder.%d.bias" % (3 * l)] = (b, (H,))
            wh, bh, K = lay.enc_wh[m], lay.enc_bh[m], (H if He >= 1 else D)
            out[e + "class_mu.weight"] = (wh, (L, K))
            out[e + "class_mu.bias"] = (bh, (L,))
            out[e + "class_logvar.weight"] = (wh + L * K, (L, K))
            out[e + "class_logvar.bias"] = (bh + L, (L,))
            if S > 0:
                out[e + "style_mu.weight"] = (wh + 2 * L * K, (S, K))
                out[e + "style_mu.bias"] = (bh + 2 * L, (S,))
                out[e + "style_logvar.weight"] = (wh + (2 * L + S) * K, (S, K))
                out[e + "style_logvar.bias"] = (bh + 2 * L + S, (S,))
        for m, name in enumerate(self.mod_names):
            D, S = self.dims[m], self.style_dims[m]
            d = "decoders.%s." % name
            K = H if Hd >= 1 else S + L
            if not self.learn_output_sample_scale:
                out[d + "logvar"] = (lay.dec_lv[m], (1, D))
            for l in range(Hd):
                out[d + "shared_decoder.%d.weight" % (3 * l)] = (lay.dec_hw[m][l], (H, S + L if l == 0 else H))
                out[d + "shared_decoder.%d.bias" % (3 * l)] = (lay.dec_hb[m][l], (H,))
            out[d + "out_mu.weight"] = (lay.dec_w[m], (D, K))
            out[d + "out_mu.bias"] = (lay.dec_b[m], (D,))
            if self.learn_output_sample_scale:
                out[d + "logvar.weight"] = (lay.dec_lvw[m], (D, K))
                out[d + "logvar.bias"] = (lay.dec_lvb[m], (D,))
        return out

    def modality_of_param(self, name):
        return self.mod_names.index(name.split(".")[1])
