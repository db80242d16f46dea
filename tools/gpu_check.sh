#!/bin/bash
# One GPU-box round: smoke, GPU parity tests, short benches.  Logs -> gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
run_bench () {   # name, args...
  local name=$1; shift
  timeout 900 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; local rc=$?
  python - "$name" "$rc" <<'PY'
import json, sys
name, rc = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/bench_{name}.json"))
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    c = d.get("cpu_baseline") or {}
    print(f"{name}: rc={rc} value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r.get('k1_ms', 0):.4f} "
          f"frac={r.get('frac', 0):.3f} e2e={e.get('value', 0):.4g} cpu={c.get('value', 0):.4g} clocks={d.get('clocks')}")
except Exception as ex:
    print(f"{name}: rc={rc} unreadable ({ex})")
    print(open(f"gpurun_out/bench_{name}.err").read()[-1500:])
PY
}
S="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
run_bench c4_hot_B20 --config c4 --batch 1048576 $S
run_bench c4_hot_B21 --config c4 --batch 2097152 $S
run_bench c4_hot_B22 --config c4 --batch 4194304 $S
run_bench c4_nohot_B21 --config c4 --batch 2097152 --no-hot $S
run_bench c4_det_B20 --config c4 --mode deterministic --batch 1048576 $S
run_bench c4u_B21 --config c4u --batch 2097152 $S
run_bench c4u_det_B20 --config c4u --mode deterministic --batch 1048576 $S
run_bench c3_B21 --config c3 --batch 2097152 $S
run_bench c5_B20 --config c5 --batch 1048576 $S
run_bench c2_B16 --config c2 --batch 65536 $S
echo "== default bench line (with cpu baseline) and reference arm"
run_bench default
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
