#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for cfg in "8 4 6" "8 2 3" "16 3 6" "32 4 6" "128 2 3"; do set -- $cfg
  for rep in 1; do echo -n "d=$1 bst=$2 ring=$3 rep=$rep: "; MFCD_K5_BSTAGES=$2 MFCD_K5_RING=$3 timeout 120 python tools/k5_stress.py --d $1 --launches 60 2>&1 | tail -1; done
done
echo "== K5"
k5 () { MFCD_K5_BSTAGES=$2 MFCD_K5_RING=$3 timeout 200 python tools/bench_k5.py --d $1 --engines tc --iters 10 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); t=d['tc']; print('d=%d bst=$2 ring=$3 ms=%.4f frac=%.3f flag=%d'%(d['d'],t['ms'],t['frac_of_hbm_peak'],t['timeout_flag']))"; }
for cfg in "8 2 6" "8 4 4" "8 3 4" "32 2 6" "32 4 4" "32 3 4" "64 3 4" "64 2 4" "96 2 4" "128 2 2" ; do k5 $cfg; done
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
