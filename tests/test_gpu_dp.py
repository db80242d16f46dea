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


def _run_worker(script, nproc, env_extra, timeout=560):
    port = 29600 + (os.getpid() * 7 + nproc) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.timeout(600)
def test_dp_through_train_model_on_one_gpu_over_gloo():
    """trainer.train_model(world_size=2) with both ranks on cuda:0 and gloo carrying the all-reduce of the
    CUDA-computed gradient: DataParallel / dp_epoch / sharded evaluation vs the oracle on a ONE-GPU box."""
    out = _run_worker("dp_api_worker.py", 2, {"MFCD_DP_TEST": "gloo1"})
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.timeout(600)
def test_dp_through_train_model_two_gpus_fused_exchange():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = _run_worker("dp_api_worker.py", 2, {"MFCD_DP_TEST": "nccl"})
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
