"""torch-CPU restatement of the reference MoPoE-VAE forward / ELBO / Adam step.

TEST INFRASTRUCTURE (see oracle/__init__.py): this is the checker the CUDA path is compared with
and the timed CPU baseline ("port") of bench.py; the product never imports it.

Every function cites the reference lines it restates (paths relative to
/root/reference/experiments).  Pinned against the unmodified reference modules executed in the
build container: tests/test_oracle_vs_reference.py and the vectors in tests/golden/ written by
oracle/make_golden.py.

Conventions shared with the C-ABI (include/mopoe_b200.h):
  * parameters: dict name -> tensor with the reference's state-dict names/shapes
    (`encoders.<m>.shared_encoder.0.weight` ... `decoders.<m>.out_mu.bias`).
  * noise: one tensor `eps[n_pass, N, E]`, E = L + sum_m S_m; columns [0:L] feed the joint
    content reparameterisation, columns [L+off_m : L+off_m+S_m] the style of modality m
    (BaseMMVae.py:143,158-159 call order: joint first, then styles in modality order).
    pass 0 is the joint pass; pass 1+m is the extra unimodal forward of modality m that
    method="poe" runs (run_epochs.py:115-117).
"""
from dataclasses import dataclass, field
from itertools import chain, combinations
from typing import List, Optional, Sequence
import math

import torch

HIDDEN = 256  # networks.py:15,50 hard-code the hidden width
POE_EPS = 1e-8  # mm_div.py:13


@dataclass
class ModelSpec:
    dims: Sequence[int] = (7, 444)
    style_dims: Sequence[int] = (3, 20)      # already zeroed when not factorised (workflow.py:148-149)
    latent_dim: int = 20
    method: str = "joint_elbo"               # poe | moe | joint_elbo | jsd  (BaseMMVae.py:43-61)
    mod_names: Sequence[str] = ("clinical", "rois")
    learn_output_scale: bool = True
    beta: float = 1.0
    beta_style: float = 1.0
    beta_content: float = 1.0
    initial_out_logvar: float = -3.0
    n_hidden_enc: int = 1                    # flags.num_hidden_layer_encoder (networks.py:16-20)
    n_hidden_dec: int = 0                    # flags.num_hidden_layer_decoder (networks.py:51-55)
    sample_scale: bool = False               # flags.learn_output_sample_scale (networks.py:58-59,73-74)
    likelihood: str = "normal"               # normal | laplace (modalities/modality.py:18-30)

    def __post_init__(self):
        self.dims = list(self.dims)
        self.style_dims = list(self.style_dims)
        self.mod_names = list(self.mod_names)[: len(self.dims)]
        assert len(self.style_dims) == len(self.dims) == len(self.mod_names)
        assert self.method in ("poe", "moe", "joint_elbo", "jsd")

    @property
    def n_mods(self):
        return len(self.dims)

    @property
    def eps_width(self):
        return self.latent_dim + sum(self.style_dims)

    def style_offset(self, m):
        return self.latent_dim + sum(self.style_dims[:m])

    def subsets(self):
        """Non-empty subsets in BaseExperiment.set_subsets order (BaseExperiment.py:58-79):
        itertools.combinations over the modality list; members of a subset sorted by NAME.
        Returns a list of (key, [modality indices in fusion order])."""
        out = []
        idx = list(range(self.n_mods))
        for combo in chain.from_iterable(combinations(idx, n) for n in range(1, len(idx) + 1)):
            names = sorted(self.mod_names[i] for i in combo)
            out.append(("_".join(names), [self.mod_names.index(n) for n in names]))
        return out


def param_shapes(spec: ModelSpec):
    """State-dict names and shapes (networks.py:9-28,44-64; measured in SURVEY.md section 5)."""
    shapes = {}
    L = spec.latent_dim
    for m, name in enumerate(spec.mod_names):
        D, S = spec.dims[m], spec.style_dims[m]
        e = "encoders.%s." % name
        width = D
        for l in range(spec.n_hidden_enc):                       # Sequential(Linear, ReLU, Dropout) per layer
            shapes[e + "shared_encoder.%d.weight" % (3 * l)] = (HIDDEN, width)
            shapes[e + "shared_encoder.%d.bias" % (3 * l)] = (HIDDEN,)
            width = HIDDEN
        shapes[e + "class_mu.weight"] = (L, width)
        shapes[e + "class_mu.bias"] = (L,)
        shapes[e + "class_logvar.weight"] = (L, width)
        shapes[e + "class_logvar.bias"] = (L,)
        if S > 0:
            shapes[e + "style_mu.weight"] = (S, width)
            shapes[e + "style_mu.bias"] = (S,)
            shapes[e + "style_logvar.weight"] = (S, width)
            shapes[e + "style_logvar.bias"] = (S,)
    for m, name in enumerate(spec.mod_names):
        D, S = spec.dims[m], spec.style_dims[m]
        d = "decoders.%s." % name
        if not spec.sample_scale:
            shapes[d + "logvar"] = (1, D)
        width = S + L
        for l in range(spec.n_hidden_dec):
            shapes[d + "shared_decoder.%d.weight" % (3 * l)] = (HIDDEN, width)
            shapes[d + "shared_decoder.%d.bias" % (3 * l)] = (HIDDEN,)
            width = HIDDEN
        shapes[d + "out_mu.weight"] = (D, width)
        shapes[d + "out_mu.bias"] = (D,)
        if spec.sample_scale:
            shapes[d + "logvar.weight"] = (D, width)
            shapes[d + "logvar.bias"] = (D,)
    return shapes


def init_params(spec: ModelSpec, seed=0, dtype=torch.float32):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
    and decoder logvar filled with initial_out_logvar (networks.py:61-64)."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, shape in param_shapes(spec).items():
        if name.endswith(".logvar") and name.startswith("decoders."):
            params[name] = torch.full(shape, spec.initial_out_logvar, dtype=dtype)
            continue
        if name.endswith(".weight"):
            fan_in = shape[1]
        else:
            fan_in = param_shapes(spec)[name[:-4] + "weight"][1]
        bound = 1.0 / math.sqrt(fan_in)
        params[name] = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return params


def trainable(spec: ModelSpec, name: str) -> bool:
    return spec.learn_output_scale or not (name.startswith("decoders.") and name.endswith(".logvar"))


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def encoder(params, spec, m, x):
    """Encoder.forward (networks.py:30-36), dropout p=0."""
    e = "encoders.%s." % spec.mod_names[m]
    h = x
    for l in range(spec.n_hidden_enc):
        h = torch.relu(h @ params[e + "shared_encoder.%d.weight" % (3 * l)].T + params[e + "shared_encoder.%d.bias" % (3 * l)])
    mu = h @ params[e + "class_mu.weight"].T + params[e + "class_mu.bias"]
    lv = h @ params[e + "class_logvar.weight"].T + params[e + "class_logvar.bias"]
    if spec.style_dims[m] > 0:
        smu = h @ params[e + "style_mu.weight"].T + params[e + "style_mu.bias"]
        slv = h @ params[e + "style_logvar.weight"].T + params[e + "style_logvar.bias"]
    else:
        smu = slv = None
    return smu, slv, mu, lv


def decoder(params, spec, m, z_style, z):
    """Decoder.forward (networks.py:66-77)."""
    d = "decoders.%s." % spec.mod_names[m]
    h = torch.cat((z_style, z), dim=1) if spec.style_dims[m] > 0 else z
    for l in range(spec.n_hidden_dec):
        h = torch.relu(h @ params[d + "shared_decoder.%d.weight" % (3 * l)].T + params[d + "shared_decoder.%d.bias" % (3 * l)])
    loc = h @ params[d + "out_mu.weight"].T + params[d + "out_mu.bias"]
    if spec.sample_scale:
        logvar = h @ params[d + "logvar.weight"].T + params[d + "logvar.bias"]
    else:
        logvar = params[d + "logvar"]
    return loc, (logvar * 0.5).exp()


def poe(mus, logvars):
    """divergence_measures/mm_div.py:13-20."""
    var = torch.exp(logvars) + POE_EPS
    T = 1.0 / var
    pd_mu = torch.sum(mus * T, dim=0) / torch.sum(T, dim=0)
    pd_var = 1.0 / torch.sum(T, dim=0)
    return pd_mu, torch.log(pd_var)


def selection_bounds(n_rows: int, n_comp: int) -> List[int]:
    """Row boundaries of utils/utils.py:63-85 for uniform weights, with the reference's own fp32
    expression: w = (1/float(K))*ones(K) (BaseMMVae.py:225), reweighted (BaseMMVae.py:99),
    i_end = i_start + int(floor(N * w[k])), last component takes the remainder."""
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


def mixture_select(mus, logvars):
    """utils/utils.py:63-85 with uniform weights: component k owns rows [b_k, b_{k+1})."""
    K, N = mus.shape[0], mus.shape[1]
    b = selection_bounds(N, K)
    mu = torch.cat([mus[k, b[k]:b[k + 1]] for k in range(K)])
    lv = torch.cat([logvars[k, b[k]:b[k + 1]] for k in range(K)])
    return mu, lv


def kl_std_normal(mu, logvar, norm):
    """divergence_measures/kl_div.py:7-14 (prior N(0,I))."""
    return -0.5 * torch.sum(1 - logvar.exp() - mu.pow(2) + logvar) / float(norm)


def kl_two_normals(mu0, logvar0, mu1, logvar1, norm):
    """divergence_measures/kl_div.py:7-14 with a second distribution."""
    return -0.5 * torch.sum(1 - logvar0.exp() / logvar1.exp() - (mu0 - mu1).pow(2) / logvar1.exp()
                            + logvar0 - logvar1) / float(norm)


def alpha_poe(alpha, mus, logvars):
    """divergence_measures/mm_div.py:23-35: weighted product of experts (the dynamic prior of the JSD objective)."""
    var = torch.exp(logvars) + POE_EPS
    a = alpha.view(-1, 1, 1).to(var.dtype)
    T = 1.0 / var
    pd_var = 1.0 / torch.sum(a * T, dim=0)
    pd_mu = pd_var * torch.sum(a * mus * T, dim=0)
    return pd_mu, torch.log(pd_var)


def normal_nll(x, loc, scale, norm, likelihood="normal"):
    """-Normal(loc, scale).log_prob(x).sum()/norm  (modalities/modality.py:42-45,
    torch.distributions.Normal.log_prob); likelihood="laplace": torch.distributions.Laplace.log_prob."""
    if likelihood == "laplace":
        scale = scale.expand_as(loc)
        return -((-torch.log(2 * scale) - torch.abs(x - loc) / scale).sum()) / float(norm)
    var = scale ** 2
    logp = -((x - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))
    return -(logp.sum()) / float(norm)


# --------------------------------------------------------------------------------------------
# forward  (BaseMMVae.forward / inference, BaseMMVae.py:137-239)
# --------------------------------------------------------------------------------------------
def inference(params, spec: ModelSpec, batch, sample=True, use_expert=None):
    """batch: dict modality name -> (N, D_m) tensor, present modalities only."""
    present = [m for m, n in enumerate(spec.mod_names) if n in batch]
    enc = {}
    for m, name in enumerate(spec.mod_names):                       # encode(), :167-178
        if m in present:
            smu, slv, mu, lv = encoder(params, spec, m, batch[name])
            enc[name + "_style"] = [smu, slv]
            enc[name] = [mu, lv]
        else:
            enc[name + "_style"] = [None, None]
            enc[name] = [None, None]
    N = batch[spec.mod_names[present[0]]].shape[0]
    L = spec.latent_dim
    mus, logvars, distr_subsets, mix_keys = [], [], {}, []
    for key, members in spec.subsets():                             # :190-216
        if not all(m in present for m in members):
            continue
        s_mus = torch.stack([enc[spec.mod_names[m]][0] for m in members])
        s_lvs = torch.stack([enc[spec.mod_names[m]][1] for m in members])
        if spec.method in ("moe", "jsd"):                           # moe_fusion :96-106 (jsd: :51-54)
            s_mu, s_lv = mixture_select(s_mus, s_lvs)
        else:                                                       # poe_fusion :109-122
            if spec.method == "poe" or len(members) == spec.n_mods:
                zeros = torch.zeros(1, N, L, dtype=s_mus.dtype)
                s_mus = torch.cat((s_mus, zeros), dim=0)
                s_lvs = torch.cat((s_lvs, zeros), dim=0)
            s_mu, s_lv = poe(s_mus, s_lvs)
        distr_subsets[key] = [s_mu, s_lv]
        if spec.method in ("moe", "jsd"):                           # fusion_condition_* :125-134
            cond = len(members) == 1
        elif spec.method == "poe":
            cond = len(members) == len(present)
        else:
            cond = True
        if cond:
            mus.append(s_mu)
            logvars.append(s_lv)
            mix_keys.append(key)
    mus = torch.stack(mus)
    logvars = torch.stack(logvars)
    if spec.method == "jsd":                                        # :217-223: the prior N(0, I) is one more component
        zeros = torch.zeros(1, N, L, dtype=mus.dtype)
        mus = torch.cat((mus, zeros), dim=0)
        logvars = torch.cat((logvars, zeros), dim=0)
        mix_keys.append("prior")
    K = mus.shape[0]
    weights = (1 / float(K)) * torch.ones(K)                        # :225
    if sample and use_expert is None:
        joint_mu, joint_lv = mixture_select(mus, logvars)           # :226-227
    elif use_expert is None:
        joint_mu, joint_lv = mus.mean(0), logvars.mean(0)           # :228-229
    else:
        joint_mu, joint_lv = distr_subsets[use_expert]              # :230-231
    return {"modalities": enc, "mus": mus, "logvars": logvars, "weights": weights,
            "joint": [joint_mu, joint_lv], "subsets": distr_subsets, "mix_keys": mix_keys}


def forward(params, spec: ModelSpec, batch, eps=None, sample_latents=True, use_expert=None):
    """BaseMMVae.forward (:137-165).  eps: (N, E) injected noise (see module docstring)."""
    latents = inference(params, spec, batch, sample=sample_latents, use_expert=use_expert)
    L = spec.latent_dim
    jmu, jlv = latents["joint"]
    if sample_latents:
        z = eps[:, :L] * (jlv * 0.5).exp() + jmu                    # reparameterize :37-40
    else:
        z = jmu
    mus, logvars = latents["mus"], latents["logvars"]
    K, N = mus.shape[0], mus.shape[1]
    dyn_prior = None
    if spec.method == "jsd":                                        # divergence_dynamic_prior :81-93
        w = latents["weights"]                                      # (not reweighted: already 1/K each)
        a_mu, a_lv = alpha_poe(w, mus, logvars)                     # calc_alphaJSD_modalities, mm_div.py:69-89
        ind = torch.stack([kl_two_normals(mus[k], logvars[k], a_mu, a_lv, N) for k in range(K)])
        dyn_prior = [a_mu, a_lv]
    else:
        w = latents["weights"] / latents["weights"].sum()           # divergence_static_prior :64-78
        ind = torch.stack([kl_std_normal(mus[k], logvars[k], N) for k in range(K)])  # mm_div.py:92-111
    results = {"latents": latents, "group_distr": latents["joint"],
               "joint_divergence": (w.to(ind.dtype) * ind).sum(), "individual_divs": ind,
               "dyn_prior": dyn_prior, "z": z, "z_style": {}}
    rec = {}
    for m, name in enumerate(spec.mod_names):                       # :155-163
        if name not in batch:
            continue
        smu, slv = latents["modalities"][name + "_style"]
        S = spec.style_dims[m]
        if S > 0 and sample_latents:
            o = spec.style_offset(m)
            zs = eps[:, o:o + S] * (slv * 0.5).exp() + smu
        else:
            zs = smu
        results["z_style"][name] = zs
        rec[name] = decoder(params, spec, m, zs, z)
    results["rec"] = rec
    return results


# --------------------------------------------------------------------------------------------
# ELBO  (run_epochs.basic_routine_epoch, run_epochs.py:73-135; utils.calc_elbo, utils.py:88-112)
# --------------------------------------------------------------------------------------------
def elbo(params, spec: ModelSpec, batch, eps):
    """eps: (n_pass, N, E); n_pass = 1 (moe, joint_elbo) or 1 + n_mods (poe)."""
    res = forward(params, spec, batch, eps[0])
    names = [n for n in spec.mod_names if n in batch]
    N = batch[names[0]].shape[0]
    log_probs = {n: normal_nll(batch[n], *res["rec"][n], N, spec.likelihood) for n in names}   # calc_log_probs :27-38
    klds = {k: kl_std_normal(mu, lv, N) for k, (mu, lv) in res["latents"]["subsets"].items()}  # :41-48
    klds_style = {}
    for n in names:                                                               # :51-59
        smu, slv = res["latents"]["modalities"][n + "_style"]
        if smu is not None:
            klds_style[n + "_style"] = kl_std_normal(smu, slv, N)
    jd = res["joint_divergence"]
    if spec.method in ("moe", "joint_elbo", "jsd"):                               # :95-103
        kld_style = sum((spec.beta_style * klds_style[n + "_style"] for n in names
                         if n + "_style" in klds_style), 0.0)                     # calc_style_kld :62-69
        total = sum(log_probs.values()) + spec.beta * (spec.beta_style * kld_style
                                                       + spec.beta_content * jd)
        uni = {}
    else:                                                                         # poe :104-128
        total = 0.0
        uni = {}
        w_style = 0.0
        for n in names:
            m = spec.mod_names.index(n)
            ks = klds_style.get(n + "_style", 0.0)
            r_mod = forward(params, spec, {n: batch[n]}, eps[1 + m])
            lp = normal_nll(batch[n], *r_mod["rec"][n], N, spec.likelihood)
            uni[n] = lp
            div = spec.beta_content * klds[n] + spec.beta_style * (spec.beta_style * ks)
            total = total + lp + spec.beta * div
            w_style = w_style + spec.beta_style * ks
        div = spec.beta_content * jd + spec.beta_style * w_style
        total = total + sum(log_probs.values()) + spec.beta * div
    return {"total_loss": total, "log_probs": log_probs, "klds": klds, "klds_style": klds_style,
            "joint_divergence": jd, "results": res, "log_probs_unimodal": uni}


def elbo_and_grads(params, spec, batch, eps):
    """loss terms + d(total_loss)/d(param) for every parameter (zeros for unused ones)."""
    p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    out = elbo(p, spec, batch, eps)
    names = list(p)
    grads = torch.autograd.grad(out["total_loss"], [p[k] for k in names], allow_unused=True)
    g = {k: (torch.zeros_like(p[k]) if gi is None else gi) for k, gi in zip(names, grads)}
    used = {k: gi is not None for k, gi in zip(names, grads)}
    for k in names:
        if not trainable(spec, k):
            g[k] = torch.zeros_like(p[k])
            used[k] = False
    return out, g, used


class Adam:
    """torch.optim.Adam defaults as configured by experiment.py:268-271: betas (0.9, 0.999),
    eps 1e-8, no weight decay, no amsgrad.  Parameters whose gradient is None in a step are
    skipped entirely (no moment decay, no step-count increment), as torch does."""

    def __init__(self, params, lr=0.002, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}

    def step(self, params, grads, used):
        for k in params:
            if not used[k]:
                continue
            self.t[k] += 1
            t = self.t[k]
            g = grads[k]
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            denom = self.v[k].sqrt() / math.sqrt(bc2) + self.eps
            params[k] = params[k] - (self.lr / bc1) * (self.m[k] / denom)
        return params


def train_steps(params, spec, batches, eps_list, lr=0.002):
    """Run len(batches) fwd+bwd+Adam steps (run_epochs.train :158-182); returns per-step losses."""
    params = {k: v.clone() for k, v in params.items()}
    opt = Adam(params, lr=lr)
    losses = []
    for batch, eps in zip(batches, eps_list):
        out, g, used = elbo_and_grads(params, spec, batch, eps)
        losses.append(out)
        params = opt.step(params, g, used)
    return params, opt, losses


def reference_eps_list(spec: ModelSpec, present, eps, single_forward=False):
    """Order in which the reference consumes `eps` through BaseMMVae.reparameterize during one
    basic_routine_epoch call: joint (N,L), then style (N,S_m) per present modality;
    in poe mode the same again for each unimodal forward."""
    L = spec.latent_dim
    out = []

    def one(pass_idx, mods):
        out.append(eps[pass_idx][:, :L])
        for m in mods:
            S = spec.style_dims[m]
            if S > 0:
                o = spec.style_offset(m)
                out.append(eps[pass_idx][:, o:o + S])

    one(0, present)
    if spec.method == "poe" and not single_forward:
        for m in present:
            one(1 + m, [m])
    return out
