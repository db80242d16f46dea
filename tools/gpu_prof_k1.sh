#!/bin/bash
# full ncu capture of K1 for one bench configuration.  usage: gpurun --timeout 900 -- 'ARGS="--config c4" NAME=c4 bash tools/gpu_prof_k1.sh'
set -u
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fwd_bwd -s 3 -c 1 -f -o gpurun_out/prof_k1_${NAME:-c4} python bench.py ${ARGS:---config c4} $A > gpurun_out/ncu_${NAME:-c4}.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
