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
import os as _os

# Concurrent sweeps (structure.parameter_scan(concurrency=k)) run every repetition on its own CUDA stream, and a
# repetition's training is a handful of LONG kernels (one persistent kernel per epoch, milliseconds each).  With the
# driver's default of 8 hardware work queues, streams share queues and a short kernel queued behind another stream's
# epoch kernel waits for it (measured: the caller thread's 4.6 ms preparation took 168 ms, the sweep ran 1.3x
# instead of ~8x faster).  One queue per stream, up to the hardware's 32; read by the driver when the CUDA context is
# created, so it has to be in the environment before the first CUDA call of the process.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402
from ._lib import lib, MfcdError, library_path  # noqa: F401,E402
