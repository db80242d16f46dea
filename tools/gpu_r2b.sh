#!/bin/bash
# Round 2, second GPU pass: tests, bench, K5 timing at d = 16/32/64/96/128, ncu of the epoch kernels and of K5.
# usage: gpurun --timeout 1700 -- 'bash tools/gpu_r2b.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; SECONDS=0
timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? (${SECONDS}s)"
tail -30 $O/pytest_gpu.log
echo "== K5 timing"
for d in 16 32 64 96 128; do
  timeout 300 python tools/bench_k5.py --d $d --engines tc --iters 10 > $O/k5_d$d.json 2> $O/k5_d$d.err; echo "k5 d=$d rc=$?"; cat $O/k5_d$d.json; tail -2 $O/k5_d$d.err
done
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; cat $O/bench_default.json; tail -5 $O/bench_default.err
echo "== ncu full: epoch kernels"
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_epoch_" -s 2 -c 2 -f -o $O/prof_epoch $CMD > $O/ncu_full_epoch.log 2>&1; echo "rc=$?"
echo "== ncu full: K5 d=64, d=128"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5_d64 python tools/bench_k5.py --d 64 --engines tc --iters 2 > $O/ncu_k5_64.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5_d128 python tools/bench_k5.py --d 128 --engines tc --iters 2 > $O/ncu_k5_128.log 2>&1; echo "rc=$?"
ls -la $O/*.ncu-rep 2>/dev/null
