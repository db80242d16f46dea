"""B200-native (sm_100a) hot path for comparison-data matrix factorisation.

Import name: ``mfcd_b200`` (the repo-root ``mfcd_b200.py`` maps it onto this
directory, whose on-disk name is not a valid python identifier).

Only what the hot path needs lives here:
  csrc/         hand-written CUDA kernels + the C ABI (include/mfcd_b200.h)
  _lib.py       ctypes binding of libmfcd_b200.so (fails loudly if missing)
  store.py      device-resident triplet records, loaders, ground-truth views
  trainer.py    model, optimiser glue, training / evaluation drivers
  metrics.py    reconstruction / alignment / correlation metrics (K5, K6)
  sampling.py   GPU triplet + label samplers (K7, K8), reference-exact host RNG
  dist.py       data-parallel plumbing over torch.distributed (NCCL)
The reference-facing API (same names and signatures as the reference's
``structure.py`` / ``generation_data.py``) is in the repo-root modules of those
names.
"""
from . import _lib  # noqa: F401
from ._lib import lib, MfcdError, library_path  # noqa: F401
