#!/bin/bash
# Round 2, last GPU pass: host-loader tests (wire8 now packed by the C packer), the default bench at HEAD (stagers
# prepared before the timed e2e epochs), ncu summaries of the kernels the previous capture's launch cap cut off.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_r2u.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (epoch / host loaders / train)"; timeout 600 python -m pytest tests/test_gpu_epoch.py tests/test_gpu_train.py -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_sub.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_sub.log
echo "== bench (default)"; SECONDS=0
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? (${SECONDS}s)"; cat $O/bench_default.json; tail -5 $O/bench_default.err
echo "== ncu: remaining secondary kernels"
timeout 900 ncu --set full --clock-control none -k regex:"k_eval|k_gt_eval|k_scores|k_rank_scatter|k_row_pearson|k_fwd_bwd_fix|k_fix_finish|k_dp_fused_adam_sync|k_recon_stats|k_sample_random|k_sample_margin|k_unique|k_unpack8|k_pack8|k_det_small" -c 30 -f -o /tmp/prof_secondary2 python tools/prof_kernels.py > $O/ncu_secondary2.log 2>&1; echo "rc=$?"; tail -2 $O/ncu_secondary2.log
python tools/ncu_summary.py raw /tmp/prof_secondary2.ncu-rep "." $O/r02b_ncu_secondary_kernels_2_summary.json
