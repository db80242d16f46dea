#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi_quick.sh N'   (DP equivalence at 2 ranks + the driver's bench command at N)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dp_gpu_worker.py > gpurun_out/dp_worker_w2.log 2>&1; echo "dp worker w=2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/dp_check_w2.json'))
for k,v in d.items(): print('   ',k,str(v)[:200])" 2>/dev/null || tail -20 gpurun_out/dp_worker_w2.log | cut -c1-400
show () {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    lines = [l for l in open(f + ".json") if l.startswith("{")]
    d = json.loads(lines[-1]); r = d["roofline"]; e = d.get("e2e") or {}
    print(f"{f}: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} e2e={e.get('value', 0):.4g} fmt={e.get('format')} exch={d['config'].get('dp_exchange')} clocks={d['clocks']}")
except Exception as ex:
    print(f, "unreadable", ex); print(open(f + ".err").read()[-2500:])
PY
}
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_w1.json 2> gpurun_out/scale_w1.err; show gpurun_out/scale_w1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/scale_w$N.json 2> gpurun_out/scale_w$N.err; show gpurun_out/scale_w$N
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 30 --warmup 5 --dp-backend nccl --no-e2e > gpurun_out/scale_w${N}_nccl.json 2> gpurun_out/scale_w${N}_nccl.err; show gpurun_out/scale_w${N}_nccl
