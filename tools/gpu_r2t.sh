#!/bin/bash
# Round 2, closing pass, second half (the first stopped at a bench.py scoping bug): default bench, 1-GPU lines at the
# config-3 / config-5 shapes, K5 X-stream rate per d, config-2 experiment timing, launch list of the bench command.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_r2t.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; cat $O/bench_default.json; tail -5 $O/bench_default.err
echo "== other shapes on one GPU (2^20 triplets per step)"
for cfg in c3 c5; do
  timeout 400 python bench.py --config $cfg --batch 1048576 --no-cpu-baseline --e2e-format records16 --no-extra-rooflines > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"
  python - $O/bench_$cfg.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], "value %.4g ms/step %.4f k1_ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["k1_ms"], (d.get("e2e") or {}).get("value", 0)))
except Exception as e:
    print("unreadable", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
echo "== K5 X-stream rate"
for d in 8 32 64 128; do
  timeout 300 python tools/bench_k5.py --d $d --iters 20 --engines tc > $O/k5_d$d.json 2> $O/k5_d$d.err || tail -5 $O/k5_d$d.err
  python -c "
import json; d=json.load(open('$O/k5_d$d.json')); t=d['tc']; print('d=$d tc: %.3f ms  %.0f GB/s  frac %.3f  %.1f TFLOP/s flag %d' % (t['ms'], t['x_stream_GBps'], t['frac_of_hbm_peak'], t['tflops'], t['timeout_flag']))"
done
echo "== config-2 experiment timing"
timeout 300 python tools/time_experiments.py > $O/time_experiments.json 2> $O/time_experiments.err; echo "rc=$?"; tail -c 1200 $O/time_experiments.json
echo "== ncu launch list"
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
timeout 300 $CMD > $O/ncu_plain1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_default.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
