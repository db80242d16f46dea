#!/bin/bash
# ncu evidence for the DEFAULT bench command (launch list + one full capture of the dominant kernel K1).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh'
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fwd_bwd -s 3 -c 2 -f -o gpurun_out/prof_k1_default $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
