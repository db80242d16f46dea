#!/bin/bash
# K5 parity tests + X-stream rate of both engines.  usage: gpurun --timeout 900 -- 'bash tools/gpu_k5.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_k5.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_k5.log
for d in 16 32 64; do
  timeout 300 python tools/bench_k5.py --d $d --engines tc > gpurun_out/k5_d$d.json 2> gpurun_out/k5_d$d.err || tail -5 gpurun_out/k5_d$d.err
  python -c "
import json; d=json.load(open('gpurun_out/k5_d$d.json')); t=d['tc']; print('d=$d tc: %.3f ms  %.0f GB/s  frac %.3f  %.1f TFLOP/s flag %d' % (t['ms'], t['x_stream_GBps'], t['frac_of_hbm_peak'], t['tflops'], t['timeout_flag']))"
done
CMD="python tools/bench_k5.py --d 64 --iters 2 --engines tc"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_recon_stats_tc -s 2 -c 1 -f -o gpurun_out/prof_k5_tc $CMD > gpurun_out/ncu_k5.log 2>&1
echo "ncu rc=$?"
