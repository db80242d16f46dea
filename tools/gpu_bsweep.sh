#!/bin/bash
# batch-size sweep of the training step at the config-4 table shape.  usage: gpurun --timeout 110 -- 'bash tools/gpu_bsweep.sh'
set -u
mkdir -p gpurun_out
run () { local name=$1; shift
  timeout 40 python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/bs_$name.json 2> gpurun_out/bs_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/bs_$name.json')); print('$name', '%.4g triplets/s' % d['value'], '%.4f ms/step' % d['ms_per_step'], 'k1 %.4f' % d['roofline']['k1_ms'], d['config']['scatter_mode'])" 2>/dev/null || tail -3 gpurun_out/bs_$name.err; }
run B64det --batch 64 --mode deterministic --steps 200 --warmup 10
run B4096 --batch 4096 --steps 200 --warmup 10
run B65536 --batch 65536 --steps 100 --warmup 5
run B1M --batch 1048576 --steps 30 --warmup 5
