#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/pytest_train.log
timeout 600 python bench.py --no-cpu-baseline ${BARGS:-} > gpurun_out/q_default.json 2> gpurun_out/q_default.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/q_default.json")); r=d["roofline"]; e=d["e2e"]
    print(f"default: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} frac={r['frac']:.3f} e2e={e['value']:.4g} fmt={e.get('format')} h2d={e['h2d_bytes_per_step']} others={e.get('other_formats')} clocks={d['clocks']}")
except Exception as ex:
    print("ERR", ex); print(open("gpurun_out/q_default.err").read()[-2000:])
PY
