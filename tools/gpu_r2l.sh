#!/bin/bash
set -u
for d in 8 16 32; do
  for cfg in "4 6" "4 4" "3 6" "3 4" "2 6" "2 3"; do set -- $cfg
    for rep in 1 2 3; do echo -n "d=$d bst=$1 ring=$2 rep=$rep: "; MFCD_K5_BSTAGES=$1 MFCD_K5_RING=$2 timeout 120 python tools/k5_stress.py --d $d --launches 40 2>&1 | tail -1; done
  done
done
for d in 64 128; do echo -n "d=$d default: "; timeout 120 python tools/k5_stress.py --d $d --launches 40 2>&1 | tail -1; done
