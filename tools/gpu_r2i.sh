#!/bin/bash
# Round 2, ninth GPU pass: K5 with its own B producer warp, sweep with 32 hardware queues, full tests + bench
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
echo "== K5"
for d in 8 16 32 64 96 128; do timeout 200 python tools/bench_k5.py --d $d --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; done
k5 () { MFCD_K5_BSTAGES=$2 MFCD_K5_RING=$3 timeout 200 python tools/bench_k5.py --d $1 --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d bst=$2 ring=$3 ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; }
for cfg in "64 2 5" "64 3 4" "64 4 3" "128 2 3" "128 3 2" "32 2 6" "32 3 6"; do k5 $cfg; done
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
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value %.4g ms/step %.4f k1 %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['roofline']['k1_ms'],d['e2e']['value']))
print(d['breakdown_ms_per_epoch'])
for r in d['rooflines_other']: print(r.get('kernel'), r.get('ms'), r.get('frac'))
print({k:v['value'] for k,v in d['e2e'].get('other_formats',{}).items()})
PY
tail -3 $O/bench_default.err
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5f_d64 python tools/bench_k5.py --d 64 --engines tc --iters 2 > $O/ncu_k5_64.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5f_d128 python tools/bench_k5.py --d 128 --engines tc --iters 2 > $O/ncu_k5_128.log 2>&1; echo "rc=$?"
