#!/bin/bash
# Round 2, 8-GPU pass: DP equivalence at 8 ranks, weak scaling at 4 and 8, strong scaling at 8, the config-4 job.
set -u
N=8
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show () {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    lines = [l for l in open(f + ".json") if l.startswith("{")]
    d = json.loads(lines[-1]); r = d["roofline"]; e = d.get("e2e") or {}
    print(f"{f}: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} e2e={e.get('value', 0):.4g} scaling={d['scaling']} loss={d.get('final_loss')} exch={(d.get('run_details') or d['config']).get('dp_exchange')} clocks={d['clocks']}")
except Exception as ex:
    print(f, "unreadable", ex); print(open(f + ".err").read()[-2500:])
PY
}
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/dp_gpu_worker.py > $O/dp_worker_w$N.log 2>&1; echo "dp worker rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/dp_check_w$N.json'))
for k,v in d.items(): print('   ',k,str(v)[:200])" 2>/dev/null || tail -20 $O/dp_worker_w$N.log | cut -c1-300
timeout 600 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e > $O/scale_w4.json 2> $O/scale_w4.err; show $O/scale_w4
timeout 600 $TR --nproc-per-node 8 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines > $O/scale_w8.json 2> $O/scale_w8.err; show $O/scale_w8
MFCD_DP_MULTIMEM=off timeout 600 $TR --nproc-per-node 8 --master-port 29529 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e > $O/scale_w8_p2p.json 2> $O/scale_w8_p2p.err; show $O/scale_w8_p2p
timeout 600 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e --scaling strong > $O/scale_strong_w8.json 2> $O/scale_strong_w8.err; show $O/scale_strong_w8
echo "== config 4 job"
SECONDS=0
timeout 1200 $TR --nproc-per-node 8 --master-port 29541 tools/run_config4.py --out $O/config4_job_w8.json > $O/config4_w8.log 2>&1; echo "config4 rc=$? (${SECONDS}s)"; grep "config4\]" $O/config4_w8.log | sort -u; tail -3 $O/config4_w8.log | cut -c1-1500
