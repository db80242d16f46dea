#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== deterministic engines over batch sizes"; timeout 600 python tools/det_sweep.py 2>&1 | tail -30
echo "== ncu secondary kernels (summaries only)"
timeout 900 ncu --set full --clock-control none -k regex:"k_adam|k_seg_reduce|k_seg_fixup|k_eval|k_gt_eval|k_rank_scatter|k_row_pearson|k_epoch_count|k_epoch_scatter|k_dp_fused_adam_sync|k_fwd_bwd_fix|k_fix_finish|k_btl_labels|k_sample_popularity" -c 36 -f -o /tmp/prof_secondary python tools/prof_kernels.py > $O/ncu_secondary.log 2>&1; echo "rc=$?"; tail -2 $O/ncu_secondary.log
python tools/ncu_summary.py raw /tmp/prof_secondary.ncu-rep "." $O/r02_ncu_secondary_kernels_summary.json
ls -la /tmp/prof_secondary.ncu-rep
