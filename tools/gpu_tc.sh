#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_metrics.py -m gpu -q -x -k "tensor_core or metrics_match or c5_shape" -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "tc tests rc=$?"
tail -30 gpurun_out/pytest_tc.log
