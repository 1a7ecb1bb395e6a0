"""CPU oracle for the MoPoE-VAE hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
leg may import anything from this package -- always as the checker or as the timed CPU baseline,
never as the thing shipped.  The product path (`2022_cambroise_interpret_multivae_b200`) never
imports it and fails loudly when its CUDA library is missing.

Contents
  mopoe_oracle.py  torch-CPU restatement of the reference model / ELBO / Adam step
                   (experiments/utils/BaseMMVae.py, run_epochs.py, divergence_measures/*, ...)
  daa_oracle.py    restatement of the DAA avatar sweep and association statistics
                   (experiments/workflow.py:361-537, experiments/stat_utils.py:55-79)
  philox.py        numpy restatement of the counter-based normal generator the CUDA path uses
                   in production mode (ours: the reference draws from torch's global generator)
  ref_harness.py   imports the UNMODIFIED reference from /root/reference (this container only)
  make_golden.py   runs the real reference on seeded inputs and writes tests/golden/*.npz

Pinning status: the model/ELBO/gradient half is pinned against the reference's own modules
executed here (tests/test_oracle_vs_reference.py + committed golden vectors).  The statistics half
(`make_regression`) depends on statsmodels, which is neither vendored nor installed: that piece is
a closed-form restatement cross-checked against scipy/numpy only -> "parity unpinned" for
stat_utils.make_regression.
"""
