#!/bin/bash
# K1 variants: parity tests of the training kernels, then short benches.  usage: gpurun --timeout 900 -- 'bash tools/gpu_k1.sh'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_k1.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/pytest_k1.log
run () {  # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/k1_$name.json 2> gpurun_out/k1_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/k1_{name}.json")); r = d["roofline"]
    print(f"{name:28s} value={d['value']:.4g} ms/step={d['ms_per_step']:.4f} k1_ms={r['k1_ms']:.4f} hot={d['config']['hot_item_rows_privatised']} loss={d['final_loss']:.6f}")
except Exception as ex:
    print(name, "ERR", ex); print(open(f"gpurun_out/k1_{name}.err").read()[-800:])
PY
}
run c4_share64_g MFCD_HOT_SMEM_KB=64 -- --config c4
run c4_share64_ng MFCD_HOT_SMEM_KB=64 -- --config c4 --no-group
run c4_share48_g MFCD_HOT_SMEM_KB=48 -- --config c4
run c4_share96_g MFCD_HOT_SMEM_KB=96 -- --config c4
run c4_share96_ng MFCD_HOT_SMEM_KB=96 -- --config c4 --no-group
run c4_noshare96_ng MFCD_K1_SHARE=0 MFCD_HOT_SMEM_KB=96 -- --config c4 --no-group
run c4_u4_share96_g MFCD_K1_UNR=4 MFCD_HOT_SMEM_KB=96 -- --config c4
run c4_u4_share96_ng MFCD_K1_UNR=4 MFCD_HOT_SMEM_KB=96 -- --config c4 --no-group
run c4_u4_share64_g MFCD_K1_UNR=4 MFCD_HOT_SMEM_KB=64 -- --config c4
run c4_u4_noshare96_ng MFCD_K1_UNR=4 MFCD_K1_SHARE=0 MFCD_HOT_SMEM_KB=96 -- --config c4 --no-group
run c4u_u4_g MFCD_K1_UNR=4 -- --config c4u
run c4u_u2_g MFCD_K1_UNR=2 -- --config c4u
