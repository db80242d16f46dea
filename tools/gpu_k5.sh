#!/bin/bash
set -u
mkdir -p gpurun_out
for d in 16 32 64; do
  timeout 300 python tools/bench_k5.py --d $d > gpurun_out/k5_d$d.json 2> gpurun_out/k5_d$d.err; echo "k5 d=$d rc=$?"; cat gpurun_out/k5_d$d.json; tail -2 gpurun_out/k5_d$d.err
done
timeout 300 python tools/bench_k5.py --d 128 --engines simt > gpurun_out/k5_d128.json 2>&1; cat gpurun_out/k5_d128.json
CMD="python tools/bench_k5.py --d 64 --iters 2 --engines tc"
$CMD > gpurun_out/k5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_recon_stats_tc -s 2 -c 1 -f -o gpurun_out/prof_k5_tc $CMD > gpurun_out/ncu_k5.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_k5.log
