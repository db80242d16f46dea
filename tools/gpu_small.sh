set -u
mkdir -p gpurun_out
timeout 600 python tools/time_experiments.py > gpurun_out/time_experiments.json 2> gpurun_out/time_experiments.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/time_experiments.json"))
for k,v in d.items(): print(k, round(v["total_s"],4), {a:round(b,4) for a,b in v["breakdown_s"].items()}, f'{v["train_triplets_per_s"]:.4g}')
PY
for cfg in "1000 1000 10 9375" "100 100 2 2000" "10000 5000 32 2000" "1000 1000 64 2000"; do timeout 120 python tools/small_step_probe.py $cfg | tail -1; MFCD_EPOCH_KERNEL=0 timeout 120 python tools/small_step_probe.py $cfg | tail -1; done > gpurun_out/small_probe.txt; cat gpurun_out/small_probe.txt
