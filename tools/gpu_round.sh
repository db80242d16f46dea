#!/bin/bash
# One GPU-box pass over HEAD: smoke, GPU parity tests, default bench + reference arm, then the ncu evidence
# for the default bench command (launch list + one full capture of K1).  Logs -> gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest -m gpu"; SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? (${SECONDS}s)"
tail -25 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
echo "== ncu"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/ncu_plain1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_bwd -s 3 -c 2 -f -o gpurun_out/prof_k1_default $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
