#!/bin/bash
# Round 2 multi-GPU pass.  usage: gpurun --gpus N -- 'bash tools/gpu_r2n.sh N [triplets_per_gpu]'
set -u
N=${1:-2}
TPG=${2:-1.25e8}
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
echo "== DP equivalence (K9 vs NCCL vs reference-recorded epoch), DP through the API"
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/dp_gpu_worker.py > $O/dp_worker_w$N.log 2>&1; echo "dp worker rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/dp_check_w$N.json'))
for k,v in d.items(): print('   ',k,str(v)[:200])" 2>/dev/null || tail -20 $O/dp_worker_w$N.log | cut -c1-300
MFCD_DP_TEST=nccl timeout 600 $TR --nproc-per-node 2 --master-port 29512 tests/dp_api_worker.py > $O/dp_api_w2.log 2>&1; echo "dp api worker (2 ranks, nccl) rc=$?"; tail -4 $O/dp_api_w2.log | cut -c1-300
echo "== bench"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines > $O/scale_w1.json 2> $O/scale_w1.err; show $O/scale_w1
for w in 2 4 8; do
  [ $w -gt $N ] && break
  timeout 600 $TR --nproc-per-node $w --master-port 2952$w bench.py --gpus $w --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines > $O/scale_w$w.json 2> $O/scale_w$w.err; show $O/scale_w$w
done
timeout 600 $TR --nproc-per-node $N --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e --scaling strong > $O/scale_strong_w$N.json 2> $O/scale_strong_w$N.err; show $O/scale_strong_w$N
timeout 600 $TR --nproc-per-node $N --master-port 29532 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e --dp-backend nccl > $O/scale_nccl_w$N.json 2> $O/scale_nccl_w$N.err; show $O/scale_nccl_w$N
echo "== config 4 job ($TPG triplets per GPU)"
SECONDS=0
timeout 1200 $TR --nproc-per-node $N --master-port 29541 tools/run_config4.py --triplets-per-gpu $TPG --out $O/config4_job_w$N.json > $O/config4_w$N.log 2>&1; echo "config4 rc=$? (${SECONDS}s)"; grep "config4\]" $O/config4_w$N.log; tail -3 $O/config4_w$N.log | cut -c1-1200
