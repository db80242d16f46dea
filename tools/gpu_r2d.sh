#!/bin/bash
# Round 2, fourth GPU pass: tests, K5 with 4 TMEM stages, bench, sweep concurrency timing, config 5.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
echo "== K5"
for d in 16 32 64 96 128; do timeout 200 python tools/bench_k5.py --d $d --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; done
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value %.4g ms/step %.4f k1 %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['roofline']['k1_ms'],d['e2e']['value']))
print(d['breakdown_ms_per_epoch'])
for r in d['rooflines_other']: print(r.get('kernel'), r.get('ms'), r.get('frac'))
PY
tail -3 $O/bench_default.err
echo "== sweep concurrency"; SECONDS=0
timeout 900 python tools/time_sweep.py --concurrency 4,8,16 > $O/time_sweep.json 2> $O/time_sweep.err; echo "rc=$? (${SECONDS}s)"; cat $O/time_sweep.json | head -40; tail -5 $O/time_sweep.err
echo "== config 5"; SECONDS=0
timeout 1200 python tools/run_config5.py > $O/config5.log 2> $O/config5.err; echo "rc=$? (${SECONDS}s)"; cut -c1-400 $O/config5.log; tail -5 $O/config5.err
echo "== ncu epoch"
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_epoch_" -s 2 -c 2 -f -o $O/prof_epoch3 $CMD > $O/ncu_full_epoch.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5b_d64 python tools/bench_k5.py --d 64 --engines tc --iters 2 > $O/ncu_k5_64.log 2>&1; echo "rc=$?"
