#!/bin/bash
# Round 2, closing pass on one GPU: smoke, the whole GPU parity suite, the default bench (now with the live host
# packer among the e2e formats) + the reference arm, 1-GPU lines at the config-3 / config-5 shapes, config 5 with a
# batch-scaled weight decay, the config-2 experiment timing, and the launch list of the default bench command.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_r2s.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc > $O/host_cores.txt; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA node\(s\)" >> $O/host_cores.txt; cat $O/host_cores.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
echo "== pytest -m gpu"; SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? (${SECONDS}s)"
tail -30 $O/pytest_gpu.log
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; cat $O/bench_default.json; tail -5 $O/bench_default.err
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cat $O/bench_reference.json
echo "== other shapes on one GPU (2^20 triplets per step)"
for cfg in c3 c5; do
  timeout 400 python bench.py --config $cfg --batch 1048576 --no-cpu-baseline --e2e-format records16 --no-extra-rooflines > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"
  python - $O/bench_$cfg.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "value %.4g ms/step %.4f k1_ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["k1_ms"], (d.get("e2e") or {}).get("value", 0)))
except Exception as e:
    print("unreadable", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
echo "== config 5 with the batch-scaled weight decay"
timeout 600 python tools/run_config5.py --out $O/config5_wd_scaled.json > $O/config5_wd_scaled.log 2>&1; echo "rc=$?"; tail -3 $O/config5_wd_scaled.log | cut -c1-400
echo "== config-2 experiment timing"
timeout 300 python tools/time_experiments.py > $O/time_experiments.json 2> $O/time_experiments.err; echo "rc=$?"; tail -c 1500 $O/time_experiments.json
echo "== ncu launch list"
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
timeout 300 $CMD > $O/ncu_plain1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_default.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
