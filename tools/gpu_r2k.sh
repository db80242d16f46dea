#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (dropin, train)"; timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_train.py -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_sub.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_sub.log
echo "== sweep concurrency"; SECONDS=0
timeout 900 python tools/time_sweep.py --concurrency 4,8,16 > $O/time_sweep.json 2> $O/time_sweep.err; echo "rc=$? (${SECONDS}s)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/time_sweep.json'))
print('seq', d['sequential_s'], 'per rep', d['per_repetition_s'])
for c,v in d['concurrent'].items():
    print(c, v['wall_s'], v['speedup'], v['identical_to_sequential'])
    for k,w in v['where'].items(): print('   ', k, w)
PY
tail -5 $O/time_sweep.err
echo "== K5 L2 prefetch sweep"
k5 () { MFCD_K5_L2PF=$2 MFCD_K5_PROBE=$3 timeout 200 python tools/bench_k5.py --d $1 --engines tc --iters 10 > $O/k5.out 2> $O/k5.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/k5.out')); t=d['tc']; print('d=%d l2pf=$2 probe=$3 ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))
except Exception as e:
    print('d=$1 l2pf=$2 probe=$3 FAILED'); print(open('gpurun_out/k5.err').read()[-600:])
"; }
for d in 8 16 32 64 128; do for pf in 0 4 8 16; do k5 $d $pf 0; done; done
for pf in 0 8 16; do k5 64 $pf 1; done
