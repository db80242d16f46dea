#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== concurrency probe (32 queues)"; timeout 300 python tools/concurrency_probe.py 2>&1 | tail -12
echo "== concurrency probe (8 queues)"; CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 python tools/concurrency_probe.py 2>&1 | tail -12
echo "== K5 d=16/32 errors"; timeout 200 python tools/bench_k5.py --d 16 --engines tc --iters 10 2>&1 | tail -5 | cut -c1-300
echo "== K5 probes"
for d in 8 64 128; do for pr in 0 1 2; do MFCD_K5_PROBE=$pr timeout 200 python tools/bench_k5.py --d $d --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('probe=$pr d=%d ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; done; done
