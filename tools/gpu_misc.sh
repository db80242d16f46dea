#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropin.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_misc.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/pytest_misc.log
MFCD_NO_PERSISTENT_EPOCH=1 timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider -k "epoch_runner or train_model_api or step_snap" > gpurun_out/pytest_misc2.log 2>&1; echo "tests (launch-per-step path) rc=$?"; tail -3 gpurun_out/pytest_misc2.log
for mode in 0 1; do
MFCD_NO_PERSISTENT_EPOCH=$mode timeout 900 python tools/time_experiments.py > gpurun_out/time_experiments_$mode.json 2> gpurun_out/time_experiments.err; echo "time rc=$?"
python - $mode <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/time_experiments_{sys.argv[1]}.json"))
for k,v in d.items():
    print("no_persistent=%s"%sys.argv[1], k, "total %.3fs"%v["total_s"], "train %.3fs"%v["breakdown_s"]["train_model"], "us/step %.1f"%(1e6*v["breakdown_s"]["train_model"]/(v["steps_per_epoch"]*v["config"]["epochs"])), "trip/s %.3g"%v["train_triplets_per_s"], "loss %.6f"%v["final_train_loss"])
PY
done
