"""Import the UNMODIFIED reference modules (read-only, /root/reference) for oracle pinning.

TEST INFRASTRUCTURE ONLY.  This module exists so that `oracle/make_golden.py` and the
`tests/test_oracle_vs_reference.py` checks can run the real reference code of the hot path
(`experiments/utils/BaseMMVae.py`, `experiments/run_epochs.py`, ...) in THIS container and compare
it with the CPU restatement in `oracle/mopoe_oracle.py`.  `/root/reference` does not exist on the
GPU box, so nothing on the product path, in `bench.py` or in the `-m gpu` tests may import this.

The reference needs a few third-party packages that are absent here (matplotlib, tensorboardX,
iterstrat, statsmodels, imageio, plotly, fire).  None of them is on the hot path; they are replaced
by empty stub modules.
"""
import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("MOPOE_REFERENCE_ROOT", "/root/reference")
_EXP = os.path.join(REFERENCE_ROOT, "experiments")


def available():
    return os.path.isdir(_EXP)


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None,
                               "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "matplotlib.lines",
    "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.axes_grid1", "tensorboardX", "iterstrat",
    "iterstrat.ml_stratifiers", "statsmodels", "statsmodels.api", "statsmodels.stats",
    "statsmodels.stats.anova", "imageio", "plotly", "plotly.express", "plotly.graph_objects",
    "fire", "seaborn", "nilearn", "nilearn.plotting", "nilearn.datasets", "nilearn.surface",
]


def install():
    """Put the reference on sys.path (once) behind stubs for its missing plotting/logging deps."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except Exception:
            mod = _Anything(name)
            mod.__path__ = []
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)
    if _EXP not in sys.path:
        sys.path.insert(0, _EXP)


def make_flags(input_dims=(7, 444), latent_dim=20, style_dim=(3, 20), method="joint_elbo",
               factorized_representation=True, num_hidden_layer_encoder=1,
               num_hidden_layer_decoder=0, likelihood="normal", learning_rate=0.002,
               batch_size=256, beta=1.0, dropout_rate=0.0, initial_out_logvar=-3.0,
               learn_output_scale=True, out_scale_per_subject=False, allow_missing_blocks=True):
    """The `flags` namespace exactly as experiments/workflow.py:98-149 builds it (hot-path fields)."""
    import torch
    flags = SimpleNamespace(
        dataset="hbn", datasetdir="", num_models=1, allow_missing_blocks=allow_missing_blocks,
        batch_size=batch_size, beta=beta, beta_1=0.9, beta_2=0.999, beta_content=1.0,
        beta_style=1.0, calc_nll=False, calc_prd=False, class_dim=latent_dim,
        data_multiplications=1, dim=64, div_weight=None, div_weight_uniform_content=None,
        end_epoch=1, eval_freq=25, eval_freq_fid=100,
        factorized_representation=factorized_representation,
        initial_learning_rate=learning_rate, initial_out_logvar=initial_out_logvar,
        input_dim=list(input_dims), joint_elbo=False, kl_annealing=0, include_prior_expert=False,
        learn_output_scale=learn_output_scale, learn_output_sample_scale=out_scale_per_subject,
        likelihood=likelihood, load_saved=False, method=method, model_save="model",
        modality_jsd=False, modality_moe=False, modality_poe=False,
        num_hidden_layer_encoder=num_hidden_layer_encoder,
        num_hidden_layer_decoder=num_hidden_layer_decoder, dropout_rate=dropout_rate,
        poe_unimodal_elbos=True, start_epoch=0, style_dim=list(style_dim),
        data_seed="defaults", grad_scaling=False)
    flags.device = torch.device("cpu")
    if method == "poe":
        flags.modality_poe = True
    elif method == "moe":
        flags.modality_moe = True
    elif method == "jsd":
        flags.modality_jsd = True
    elif method == "joint_elbo":
        flags.joint_elbo = True
    else:
        raise ValueError(method)
    flags.num_mods = len(flags.input_dim)
    flags.div_weight_uniform_content = 1 / (flags.num_mods + 1)
    flags.alpha_modalities = [flags.div_weight_uniform_content]
    flags.div_weight = 1 / (flags.num_mods + 1)
    flags.alpha_modalities.extend([flags.div_weight for _ in range(flags.num_mods)])
    if not flags.factorized_representation:
        flags.style_dim = [0] * len(flags.style_dim)
    return flags


_MOD_NAMES4 = ["clinical", "rois", "modc", "modd"]


def build_reference_model(flags, seed=0, mod_names=None):
    """Instantiate the reference VAE (networks/VAE.py:6-8) the way
    MultimodalExperiment.set_modalities/set_subsets/set_models do (experiment.py:123-144,
    BaseExperiment.py:58-79), but generic in M (set_modalities hard-codes two)."""
    install()
    import torch
    from itertools import chain, combinations
    from modalities.modality import Modality
    from multimodal_cohort.networks.VAE import VAE
    from multimodal_cohort.networks.networks import Encoder, Decoder

    class _Mod(Modality):
        def save_data(self, d, fn, args):
            pass

        def plot_data(self, d):
            return d

    M = len(flags.input_dim)
    names = list(mod_names or _MOD_NAMES4[:M])
    mods = {n: _Mod(n, Encoder, Decoder, flags.class_dim, flags.style_dim[m], flags.likelihood)
            for m, n in enumerate(names)}
    xs = list(mods)
    subsets = {}
    for combo in chain.from_iterable(combinations(xs, n) for n in range(len(xs) + 1)):
        subsets["_".join(sorted(combo))] = [mods[n] for n in sorted(combo)]
    torch.manual_seed(seed)
    model = VAE(flags, mods, subsets)
    exp = SimpleNamespace(flags=flags, modalities=mods, subsets=subsets, models=model,
                          rec_weights={n: 1.0 for n in names},
                          style_weights={n: flags.beta_style for n in names})
    return model, exp


class InjectedNoise:
    """Context manager: make BaseMMVae.reparameterize (BaseMMVae.py:37-40) consume a given list
    of eps tensors, in call order, instead of the global torch generator."""

    def __init__(self, eps_list):
        self.eps = list(eps_list)
        self.calls = []

    def __enter__(self):
        install()
        from utils.BaseMMVae import BaseMMVae
        self._cls = BaseMMVae
        self._orig = BaseMMVae.reparameterize
        outer = self

        def reparameterize(self, mu, logvar):
            std = logvar.mul(0.5).exp()
            eps = outer.eps.pop(0)
            assert eps.shape == std.shape, (eps.shape, std.shape)
            outer.calls.append(tuple(std.shape))
            return eps.mul(std).add(mu)

        BaseMMVae.reparameterize = reparameterize
        return self

    def __exit__(self, *a):
        self._cls.reparameterize = self._orig
        return False
