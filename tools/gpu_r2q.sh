#!/bin/bash
# Round 2: deterministic fixed-point engine (tests + timing), ncu summaries of the secondary kernels
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest (train, epoch, dropin)"; timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_epoch.py tests/test_gpu_dropin.py -m gpu -q --maxfail=20 -p no:cacheprovider > $O/pytest_sub.log 2>&1; echo "rc=$?"; tail -8 $O/pytest_sub.log
echo "== deterministic mode timing (bench --mode deterministic, 2^20 per step)"
for cfg in c4 c4u; do
  timeout 600 python bench.py --config $cfg --mode deterministic --batch 1048576 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-rooflines > $O/bench_det_$cfg.json 2> $O/bench_det_$cfg.err; python - $O/bench_det_$cfg.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], 'ms/step %.4f k1 %.4f value %.4g launches %s'%(d['ms_per_step'],d['roofline']['k1_ms'],d['value'],d['gpu_launches']))
except Exception as e: print('unreadable', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
echo "== ncu secondary kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_adam|k_seg_reduce|k_seg_fixup|k_eval|k_gt_eval|k_rank_scatter|k_row_pearson|k_epoch_count|k_epoch_scatter|k_dp_fused_adam_sync|k_fwd_bwd_fix|k_fix_finish|k_btl_labels|k_sample_popularity|k_recon_stats" -c 40 -f -o $O/prof_secondary python tools/prof_kernels.py > $O/ncu_secondary.log 2>&1; echo "rc=$?"; tail -3 $O/ncu_secondary.log
