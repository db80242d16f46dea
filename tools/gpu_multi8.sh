#!/bin/bash
# usage: gpurun --gpus 8 -- 'bash tools/gpu_multi8.sh'   (the driver's bench command at 8 and 4 GPUs + DP equivalence at 8)
set -u
mkdir -p gpurun_out
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
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/scale_w8.json 2> gpurun_out/scale_w8.err; show gpurun_out/scale_w8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 30 --warmup 5 --no-e2e > gpurun_out/scale_w4.json 2> gpurun_out/scale_w4.err; show gpurun_out/scale_w4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 tests/dp_gpu_worker.py > gpurun_out/dp_worker_w8.log 2>&1; echo "dp worker w=8 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/dp_check_w8.json'))
for k,v in d.items(): print('   ',k,str(v)[:160])" 2>/dev/null || tail -20 gpurun_out/dp_worker_w8.log | cut -c1-400
