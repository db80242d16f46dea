"""Data-parallel equivalence on real GPUs (needs >= 2 devices; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_two_gpu_dp_matches_reference_and_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=560)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
