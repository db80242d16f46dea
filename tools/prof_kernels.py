#!/usr/bin/env python
"""Launches each secondary kernel of the hot path once or twice on a realistic shape, so that one `ncu --set full`
pass (tools/gpu_r2_prof.sh) can capture them: K2 deterministic scatter (k_seg_reduce), K3 (k_adam), K4 (k_eval),
K6 (k_rank_scatter / k_row_pearson), the epoch batching kernels, K7/K8 samplers.  Not a benchmark."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import mfcd_b200  # noqa: E402
from mfcd_b200 import metrics, sampling, trainer  # noqa: E402
from mfcd_b200._lib import lib, check, ptr, current_stream  # noqa: E402
from mfcd_b200.store import GroundTruth, TripletLoader  # noqa: E402
from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch, eval_batches  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    n, m, d, B = 100_000, 50_000, 64, 1 << 20
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    A, _ = torch.linalg.qr(torch.randn(n, d, generator=gen, device=dev))
    Bm, _ = torch.linalg.qr(torch.randn(m, d, generator=gen, device=dev))
    gt = GroundTruth(A=A, B=Bm, scale=math.sqrt(n * m) / (2 * math.sqrt(d)), device=dev)
    # K7 + dedup + K8
    ts = sampling.sample_popularity(gt, 4 * B, seed=3)
    store = sampling.btl_records(gt, ts, scale=1.0, K=1, soft=False, seed=4)
    torch.manual_seed(0)
    model = MatrixFactorization(n, m, d)
    fs = model.flat_state(dev)
    spec = OptimizerSpec.adam(lr=1e-3, weight_decay=1e-5)
    # epoch batching + K1 atomic + K3
    loader = TripletLoader(store, B, shuffle=True, shuffle_rng="device")
    trainer.train_epoch(fs, loader, spec, 0)
    # deterministic mode at 2^20: the fixed-point engine (k_fwd_bwd_fix + k_fix_finish)
    run_epoch(fs, store.slice(0, 2 * B), None, B, spec, 1)
    # ... and the sort engine (K2: k_seg_reduce + fix-ups after three radix sorts)
    sub = store.slice(0, B)
    need = C.c_size_t(0)
    check(lib.mfcd_det_workspace_bytes(B, d, C.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, device=dev)
    nU = n * d
    check(lib.mfcd_triplet_fwd_bwd_det_sort(ptr(fs.params), ptr(fs.params[nU:]), ptr(sub.rec), None, 0, B, d, 1.0 / B, n, m,
                                            ptr(fs.grads), ptr(fs.grads[nU:]), ptr(loss), ptr(ws), need.value,
                                            current_stream()), "det_sort")
    fs.grads.zero_()
    # K9 on one rank (the rank exchanges with itself through the same flag protocol)
    flags = torch.zeros(64, dtype=torch.int32, device=dev); counter = torch.zeros(1, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev); other = torch.zeros_like(fs.grads)
    arr = lambda t: (C.c_uint64 * 1)(t.data_ptr())
    for seq in (1, 2):
        check(lib.mfcd_dp_fused_adam_sync(arr(fs.grads), arr(fs.params), arr(flags), 0, 0, 0, 1, fs.params.numel(),
                                          ptr(fs.state1), ptr(fs.state2), 1e-3, 0.9, 0.999, 1e-8, 1e-5, seq, seq, ptr(counter),
                                          ptr(err), ptr(other), current_stream()), "k9")
    # K4
    eval_batches(fs, store, 64)
    # K5 + K6 on a dense block
    xn, xm = 4096, 20480
    km = MatrixFactorization(xn, xm, d)
    g5 = GroundTruth(X=torch.randn(xn, xm, device=dev), device=dev)
    metrics._row_stats_device(km, g5, 1.0)
    metrics.row_spearman(km, g5)
    torch.cuda.synchronize()
    print("prof_kernels ok")


if __name__ == "__main__":
    main()
