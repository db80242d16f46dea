import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, so a bare
    `pytest tests/` works in the CPU container."""
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def batches_from(g):
    """Rebuild the exact batch list the reference trained on."""
    sizes = g["batch_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    out = []
    for k in range(len(sizes)):
        a, b = offs[k], offs[k + 1]
        out.append((g["batch_u"][a:b], g["batch_i"][a:b], g["batch_j"][a:b], g["batch_z"][a:b]))
    return out
