#!/bin/bash
# Round 2: throughput at the config-3 and config-5 table shapes on 2, 4 and 8 GPUs (weak scaling, 2^20 triplets per
# GPU per step).  The 2- and 4-GPU runs use disjoint GPUs of the same box at the same time (device-timed, so they do
# not disturb each other's numbers beyond sharing the host and the switch); the 8-GPU run has the box to itself.
# usage: gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_r2v.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ARGS="--batch 1048576 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines"
show () {
  python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f + ".json") if l.startswith("{")][-1]); r = d["roofline"]
    print(f"{f}: value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} loss={d.get('final_loss')} exch={(d.get('run_details') or {}).get('dp_exchange')} clocks={d['clocks']}")
except Exception as ex:
    print(f, "unreadable", ex); print(open(f + ".err").read()[-1500:])
PY
}
for cfg in c3 c5; do
  CUDA_VISIBLE_DEVICES=0,1 timeout 300 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --config $cfg $ARGS > $O/scale_${cfg}_w2.json 2> $O/scale_${cfg}_w2.err &
  P2=$!
  CUDA_VISIBLE_DEVICES=2,3,4,5 timeout 300 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --config $cfg $ARGS > $O/scale_${cfg}_w4.json 2> $O/scale_${cfg}_w4.err &
  P4=$!
  CUDA_VISIBLE_DEVICES=6 timeout 300 python bench.py --gpus 1 --config $cfg $ARGS > $O/scale_${cfg}_w1.json 2> $O/scale_${cfg}_w1.err &
  P1=$!
  wait $P2 $P4 $P1
  show $O/scale_${cfg}_w1; show $O/scale_${cfg}_w2; show $O/scale_${cfg}_w4
  timeout 300 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --config $cfg $ARGS > $O/scale_${cfg}_w8.json 2> $O/scale_${cfg}_w8.err
  show $O/scale_${cfg}_w8
done
