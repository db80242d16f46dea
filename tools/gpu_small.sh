set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropin.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_small.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/pytest_small.log
for cfg in "1000 1000 10 9375" "100 100 2 2000" "10000 5000 32 2000" "1000 1000 64 2000" "300 300 16 2000"; do timeout 120 python tools/small_step_probe.py $cfg | tail -1; MFCD_EPOCH_CLUSTER=0 timeout 120 python tools/small_step_probe.py $cfg | tail -1; done
