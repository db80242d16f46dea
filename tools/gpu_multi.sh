#!/bin/bash
# Multi-GPU round: DP equivalence worker + scaling bench for both exchange back ends.
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "== DP equivalence"
for w in 2 $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 2951$w tests/dp_gpu_worker.py > gpurun_out/dp_worker_w$w.log 2>&1; echo "dp worker w=$w rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/dp_check_w$w.json'))
for k,v in d.items(): print('   ',k,str(v)[:300])" 2>/dev/null || tail -20 gpurun_out/dp_worker_w$w.log | cut -c1-400
  [ "$N" = "2" ] && break
done
show () {  # file
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    lines = [l for l in open(f + ".json") if l.startswith("{")]
    d = json.loads(lines[-1]); r = d["roofline"]; e = d.get("e2e") or {}
    print(f"{f}: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} host_ms={d.get('host_wall_ms_per_step', 0):.4f} "
          f"k1_ms={r['k1_ms']:.4f} e2e={e.get('value', 0):.4g} exch={d['config'].get('dp_exchange')} clocks={d['clocks']}")
except Exception as ex:
    print(f, "unreadable", ex); print(open(f + ".err").read()[-2500:])
PY
}
echo "== scaling bench"
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_w1.json 2> gpurun_out/scale_w1.err; show gpurun_out/scale_w1
for w in 2 4 8; do
  [ $w -gt $N ] && break
  for be in peer nccl; do
    extra=""; [ $be = peer ] && extra="--no-e2e"; [ $be = nccl ] && extra="--no-e2e"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 2952$w bench.py --gpus $w --steps 30 --warmup 5 --dp-backend $be --multimem on $extra > gpurun_out/scale_w${w}_$be.json 2> gpurun_out/scale_w${w}_$be.err
    show gpurun_out/scale_w${w}_$be
  done
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 2953$w bench.py --gpus $w --steps 30 --warmup 5 --dp-backend peer --multimem off --no-e2e > gpurun_out/scale_w${w}_p2p.json 2> gpurun_out/scale_w${w}_p2p.err
  show gpurun_out/scale_w${w}_p2p
  # the driver's form of the command (default flags, with e2e)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port 2954$w bench.py --gpus $w --steps 30 --warmup 5 > gpurun_out/scale_w${w}.json 2> gpurun_out/scale_w${w}.err
  show gpurun_out/scale_w${w}
done
