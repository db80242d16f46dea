#!/bin/bash
# Round 2, fifth GPU pass: tests after the concurrency / fixed-point eval fixes, K1 span kernel vs lean, sweep timing.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
Q="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
show () { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print('value %.4g ms/step %.4f k1 %.4f loss %.6f'%(d['value'],d['ms_per_step'],d['roofline']['k1_ms'],d.get('final_loss',0)), d.get('breakdown_ms_per_epoch'))
PY
}
echo "== bench span (default)"; timeout 600 python bench.py $Q > $O/bench_span.json 2> $O/bench_span.err; echo "rc=$?"; show $O/bench_span.json
echo "== bench lean (MFCD_K1_SPAN=0)"; MFCD_K1_SPAN=0 timeout 600 python bench.py $Q > $O/bench_lean.json 2> $O/bench_lean.err; echo "rc=$?"; show $O/bench_lean.json
for mt in 1 2 8 16; do echo "== span min tiles $mt"; MFCD_K1_SPAN_MIN_TILES=$mt timeout 600 python bench.py $Q > $O/bench_span_mt$mt.json 2>/dev/null; show $O/bench_span_mt$mt.json; done
echo "== bench uniform items"; for sp in 1 0; do MFCD_K1_SPAN=$sp timeout 600 python bench.py $Q --config c4u > $O/bench_uni_$sp.json 2>/dev/null; show $O/bench_uni_$sp.json; done
echo "== sweep concurrency"; SECONDS=0
timeout 900 python tools/time_sweep.py --concurrency 4,8,16 > $O/time_sweep.json 2> $O/time_sweep.err; echo "rc=$? (${SECONDS}s)"; cat $O/time_sweep.json | head -40; tail -5 $O/time_sweep.err
echo "== ncu span"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_bwd_span" -s 5 -c 1 -f -o $O/prof_k1_span python bench.py $Q > $O/ncu_k1_span.log 2>&1; echo "rc=$?"
