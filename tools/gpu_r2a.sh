#!/bin/bash
# Round 2, first GPU pass: smoke, GPU parity tests, bench through the product API (+ the no-shuffle variant and the
# reference arm), launch list of the default bench command, full ncu captures of K1 and the epoch kernels.
# usage: gpurun --timeout 1700 -- 'bash tools/gpu_r2a.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
echo "== pytest -m gpu"; SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? (${SECONDS}s)"
tail -40 $O/pytest_gpu.log
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; cat $O/bench_default.json; tail -5 $O/bench_default.err
echo "== bench (no shuffle)"
timeout 600 python bench.py --no-shuffle --no-cpu-baseline --no-e2e --no-extra-rooflines > $O/bench_noshuffle.json 2> $O/bench_noshuffle.err; echo "rc=$?"; cat $O/bench_noshuffle.json
echo "== reference arm"
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cat $O/bench_reference.json
echo "== ncu launch list"
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
timeout 300 $CMD > $O/ncu_plain1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_default.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
echo "== ncu full: K1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_bwd -s 3 -c 1 -f -o $O/prof_k1_default $CMD > $O/ncu_full_k1.log 2>&1; echo "rc=$?"
echo "== ncu full: epoch kernels + adam"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_epoch_|k_adam" -s 3 -c 3 -f -o $O/prof_epoch_adam $CMD > $O/ncu_full_epoch.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep 2>/dev/null
