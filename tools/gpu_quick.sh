#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "train tests rc=$?"; tail -4 gpurun_out/pytest_train.log
S="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
for cfg in "c4 4194304" "c4u 4194304" "c3 2097152" "c5 1048576" "c2 65536"; do
  set -- $cfg
  timeout 300 python bench.py --config $1 --batch $2 $S > gpurun_out/q_$1.json 2> gpurun_out/q_$1.err
  python - $1 <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/q_{sys.argv[1]}.json")); r=d["roofline"]
    print(f"{sys.argv[1]}: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} frac={r['frac']:.3f}")
except Exception as e:
    print(sys.argv[1],"ERR",e); print(open(f"gpurun_out/q_{sys.argv[1]}.err").read()[-1500:])
PY
done
