"""Process-wide switches of the hot path (read once from the environment,
changeable at run time through ``set_rng_mode`` / ``set_scatter_mode``)."""
import os

# "device":    GPU samplers / label draws / epoch shuffles (Philox streams seeded from torch's global seed)
# "reference": replay torch / numpy HOST generators exactly as the reference consumes them, so a seeded
#              run sees the reference's triplets, labels, split and batch order (parity runs, small sizes)
RNG_MODE = os.environ.get("MFCD_RNG", "device")
# "auto" | "atomic" | "deterministic"  (see trainer.resolve_mode)
SCATTER_MODE = os.environ.get("MFCD_MODE", "auto")
# dense ground-truth matrices above this many elements are kept as low-rank factors
DENSE_X_MAX_ELEMS = int(os.environ.get("MFCD_DENSE_X_MAX", str(1 << 31)))


def set_rng_mode(mode):
    global RNG_MODE
    if mode not in ("device", "reference"):
        raise ValueError("rng mode must be 'device' or 'reference'")
    RNG_MODE = mode


def set_scatter_mode(mode):
    global SCATTER_MODE
    if mode not in ("auto", "atomic", "deterministic"):
        raise ValueError("scatter mode must be 'auto', 'atomic' or 'deterministic'")
    SCATTER_MODE = mode
