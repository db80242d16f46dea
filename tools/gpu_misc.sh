#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 900 python tools/time_experiments.py > gpurun_out/time_experiments.json 2> gpurun_out/time_experiments.err; echo "time rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/time_experiments.json"))
for k,v in d.items():
    print(k, "total %.3fs"%v["total_s"], "train %.3fs"%v["breakdown_s"]["train_model"], "us/step %.1f"%(1e6*v["breakdown_s"]["train_model"]/(v["steps_per_epoch"]*v["config"]["epochs"])), "trip/s %.3g"%v["train_triplets_per_s"], "loss %.6f"%v["final_train_loss"])
PY
