#!/bin/bash
set -u
N=${1:-2}
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
    print(f"{f}: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} e2e={e.get('value', 0):.4g} scaling={d['scaling']} loss={d.get('final_loss')} exch={(d.get('run_details') or d['config']).get('dp_exchange')}")
except Exception as ex:
    print(f, "unreadable", ex); print(open(f + ".err").read()[-2500:])
PY
}
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/dp_gpu_worker.py > $O/dp_worker_w$N.log 2>&1; echo "dp worker rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/dp_check_w$N.json'))
print({k:v for k,v in d.items() if k in ('ok','peer_ok','replicas_identical','golden_loss_rel','peer_p2p_params_rel','peer_multimem_params_rel')})" 2>/dev/null || tail -20 $O/dp_worker_w$N.log | cut -c1-300
MFCD_DP_TEST=nccl timeout 600 $TR --nproc-per-node 2 --master-port 29512 tests/dp_api_worker.py > $O/dp_api_w2.log 2>&1; echo "dp api worker (2 ranks, nccl) rc=$?"; tail -1 $O/dp_api_w2.log | cut -c1-400
for mm in auto; do
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e > $O/scale_db_w$N.json 2> $O/scale_db_w$N.err; show $O/scale_db_w$N
MFCD_DP_DOUBLE_BUFFER=0 timeout 600 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extra-rooflines --no-e2e > $O/scale_sb_w$N.json 2> $O/scale_sb_w$N.err; show $O/scale_sb_w$N
done
