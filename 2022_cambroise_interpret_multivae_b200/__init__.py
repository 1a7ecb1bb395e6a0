"""B200-native (sm_100a) implementation of the MoPoE-VAE hot path of
neurospin-projects/2022_cambroise_interpret_multivae: the joint-ELBO training step and the Digital
Avatars Analysis sweep, behind the reference's own Python interfaces.

  spec.PathSpec            flags -> model description / subset table / selection boundaries
  engine                   functional wrappers of the C-ABI (include/mopoe_b200.h)
  model.VAE                drop-in for multimodal_cohort.networks.VAE.VAE (same state-dict keys,
                           forward(input_batch, sample_latents, use_expert) -> results dict)
  run_epochs               basic_routine_epoch / train / test / run_epochs mirrors
  daa                      daa_sweep / daa_regression / sharding helpers
  workflow                 train_exp / daa_exp mirrors (CLI contract)
  stat_utils               make_regression mirror

The package directory name starts with a digit, so import it through the `mopoe_b200` alias module
at the repository root (or importlib.import_module).  There is NO CPU fallback: every compute entry
point raises when libmopoe_b200.so or a CUDA device is missing.
"""
from . import _lib  # noqa: F401
from ._lib import MopoeError, build  # noqa: F401
from .spec import PathSpec, selection_bounds  # noqa: F401
