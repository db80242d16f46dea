#!/bin/bash
# Round 2, eighth GPU pass: K5 two-group epilogue, K1 span v3 (packed), host profile of the sweep preparation
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (train, metrics)"; timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_metrics.py tests/test_gpu_epoch.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_sub.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_sub.log
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
show () { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('value %.4g ms/step %.4f k1 %.4f loss %.6f'%(d['value'],d['ms_per_step'],d['roofline']['k1_ms'],d.get('final_loss',0)), d.get('breakdown_ms_per_epoch'))
PY
}
echo "== bench span v3"; timeout 600 python bench.py $Q > $O/bench_span.json 2> $O/bench_span.err; echo "rc=$?"; show $O/bench_span.json
echo "== bench uniform"; timeout 600 python bench.py $Q --config c4u > $O/bench_uni.json 2>/dev/null; show $O/bench_uni.json
echo "== bench c5 shape (d=128)"; timeout 600 python bench.py $Q --config c5 > $O/bench_c5.json 2>/dev/null; show $O/bench_c5.json
echo "== K5"
for d in 8 16 32 64 96 128; do timeout 200 python tools/bench_k5.py --d $d --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; done
echo "== host profile"; timeout 600 python tools/prof_prepare.py > $O/prof_prepare.txt 2>&1; echo "rc=$?"; head -120 $O/prof_prepare.txt | cut -c1-180
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_bwd_span" -s 5 -c 1 -f -o $O/prof_k1_span3 python bench.py $Q > $O/ncu_k1_span.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_recon_stats_tc" -s 3 -c 1 -f -o $O/prof_k5e_d64 python tools/bench_k5.py --d 64 --engines tc --iters 2 > $O/ncu_k5_64.log 2>&1; echo "rc=$?"
