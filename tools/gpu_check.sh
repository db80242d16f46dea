#!/bin/bash
# One GPU-box round: smoke, GPU parity tests, short benches.  Logs -> gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_gpu.log
for spec in "c4u atomic 1048576" "c4 atomic 1048576" "c4 deterministic 1048576" "c4u deterministic 1048576"; do
  set -- $spec
  echo "== bench $1 $2 B=$3"
  timeout 600 python bench.py --config $1 --mode $2 --batch $3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err; echo "rc=$?"
  cat gpurun_out/bench_$1_$2.json; tail -3 gpurun_out/bench_$1_$2.err
done
