#!/bin/bash
# K1 tuning experiments (variants of in-flight depth / occupancy / hot smem budget)
set -u
mkdir -p gpurun_out
S="--steps 20 --warmup 3 --no-cpu-baseline --no-e2e --batch 2097152"
for var in 0 2; do
 for cfg in c4 c4u; do
  for kb in 64 48; do
   if [ $cfg = c4u ] && [ $kb = 48 ]; then continue; fi
   MFCD_K1_VARIANT=$var MFCD_HOT_SMEM_KB=$kb timeout 300 python bench.py --config $cfg $S > gpurun_out/var_${cfg}_v${var}_kb${kb}.json 2>> gpurun_out/var.err
   python - <<PY
import json
d=json.load(open("gpurun_out/var_${cfg}_v${var}_kb${kb}.json")); r=d["roofline"]
print("cfg=$cfg var=$var kb=$kb value=%.4g k1_ms=%.4f hot=%s clocks=%s" % (d["value"], r["k1_ms"], d["config"].get("hot_item_rows_privatised"), d["clocks"]))
PY
  done
 done
done
for var in 0 2; do
  MFCD_K1_VARIANT=$var timeout 300 python bench.py --config c5 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --batch 1048576 > gpurun_out/var_c5_v$var.json 2>>gpurun_out/var.err
  MFCD_K1_VARIANT=$var timeout 300 python bench.py --config c3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --batch 2097152 > gpurun_out/var_c3_v$var.json 2>>gpurun_out/var.err
  python - <<PY
import json
for c in ("c5","c3"):
    d=json.load(open("gpurun_out/var_%s_v$var.json"%c)); print(c,"var=$var value=%.4g k1_ms=%.4f"%(d["value"],d["roofline"]["k1_ms"]))
PY
done
