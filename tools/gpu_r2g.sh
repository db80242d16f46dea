#!/bin/bash
# Round 2, seventh GPU pass: K5 with elected issuers, sweep trace
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (metrics, dropin)"; timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_dropin.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_sub.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_sub.log
echo "== K5"
for d in 8 16 32 64 96 128; do timeout 200 python tools/bench_k5.py --d $d --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; done
echo "== K5 stage/ring sweep"
k5 () { MFCD_K5_BSTAGES=$2 MFCD_K5_RING=$3 timeout 200 python tools/bench_k5.py --d $1 --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d bst=$2 ring=$3 ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; }
for cfg in "64 2 5" "64 2 4" "64 3 4" "128 2 3" "128 2 2" "128 3 2" "32 2 6" "32 4 5"; do k5 $cfg; done
echo "== sweep concurrency"; SECONDS=0
timeout 900 python tools/time_sweep.py --concurrency 8 > $O/time_sweep.json 2> $O/time_sweep.err; echo "rc=$? (${SECONDS}s)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/time_sweep.json'))
print('seq', d['sequential_s'], 'per rep', d['per_repetition_s'])
for c,v in d['concurrent'].items():
    print(c, v['wall_s'], v['speedup'], v['identical_to_sequential'])
    for k,w in v['where'].items(): print('   ', k, w)
PY
tail -5 $O/time_sweep.err
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5d_d64 python tools/bench_k5.py --d 64 --engines tc --iters 2 > $O/ncu_k5_64.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5d_d128 python tools/bench_k5.py --d 128 --engines tc --iters 2 > $O/ncu_k5_128.log 2>&1; echo "rc=$?"
