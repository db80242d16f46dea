#!/bin/bash
# Round 2, third GPU pass: K5 stage/ring sweep, epoch-batching timing, tests, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (epoch + metrics)"; timeout 600 python -m pytest tests/test_gpu_epoch.py tests/test_gpu_metrics.py -m gpu -q -p no:cacheprovider > $O/pytest_sub.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_sub.log
echo "== K5 sweep"
k5 () { # d bst ring
  MFCD_K5_BSTAGES=$2 MFCD_K5_RING=$3 timeout 200 python tools/bench_k5.py --d $1 --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d bst=$2 ring=$3 ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"
}
for cfg in "16 2 6" "16 4 5" "16 4 6" "32 2 6" "32 3 5" "32 4 5" "64 2 5" "64 3 4" "64 4 3" "96 2 4" "96 3 2" "128 2 3"; do k5 $cfg; done
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('value %.4g ms/step %.4f k1 %.4f e2e %.4g'%(d['value'],d['ms_per_step'],d['roofline']['k1_ms'],d['e2e']['value']))
print(d['breakdown_ms_per_epoch'])
for r in d['rooflines_other']: print(r.get('kernel'), r.get('ms'), r.get('frac'))
print({k:v['value'] for k,v in d['e2e']['other_formats'].items()})
PY
tail -3 $O/bench_default.err
echo "== full pytest"; timeout 900 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
echo "== ncu full: epoch kernels"
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_epoch_" -s 2 -c 2 -f -o $O/prof_epoch2 $CMD > $O/ncu_full_epoch.log 2>&1; echo "rc=$?"
