"""Properties at BASELINE.json's full table shapes (configs 3-5) that do not need the CPU oracle:
sampler predicates and idempotence, label statistics, evaluation consistency."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


def test_c3_margin_sampler_at_scale(G):
    """config 3 (10 000 x 5 000, d = 32, margin strategy): 2M unique close-call triplets from a factored X."""
    from generation_data import generate_low_rank_gpu
    from mfcd_b200 import sampling
    from mfcd_b200._lib import lib, check, ptr, current_stream
    n, m, N = 10_000, 5_000, 2_000_000
    torch.manual_seed(3)
    gt = generate_low_rank_gpu(n, m, 32, "cuda", seed=3, force_factored=True)
    ts = sampling.sample_margin(gt, N, None, max_attempts=40_000_000)
    assert len(ts) == N
    u, i, j = ts.columns()
    assert (i != j).all() and int(u.max()) < n and int(torch.maximum(i, j).max()) < m
    assert torch.unique(ts.keys).numel() == N                                  # uniqueness
    thr = sampling.margin_threshold(gt, N)
    A, B = gt.A, gt.B
    diff = ((A[u] * B[i]).sum(1) - (A[u] * B[j]).sum(1)) * gt.scale
    assert (diff.abs() <= thr * (1 + 1e-4) + 1e-5).all()                       # the margin predicate holds
    # idempotence of the dedup: feeding the accepted keys again with themselves as `seen` accepts nothing
    need = C.c_size_t(0)
    check(lib.mfcd_unique_workspace_bytes(N, N, C.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=G.DEV)
    out = torch.empty(N, dtype=torch.int64, device=G.DEV); n_out = torch.zeros(1, dtype=torch.int64, device=G.DEV)
    check(lib.mfcd_unique_accept(ptr(ts.keys), N, ptr(ts.keys), N, N, ptr(out), ptr(n_out), ptr(ws), need.value,
                                 current_stream()), "uniq")
    assert int(n_out.item()) == 0


def test_c4_popularity_labels_and_eval_consistency(G):
    """config 4 shape (100 000 x 50 000, d = 64, zipf items, factored X): label frequency tracks the BTL
    probability, ground-truth accuracy beats chance, and K4's loss equals K1's loss on the same records."""
    from generation_data import generate_low_rank_gpu
    from mfcd_b200 import sampling
    from mfcd_b200._lib import lib, check, ptr, current_stream
    from mfcd_b200.store import TripletLoader
    from mfcd_b200.trainer import MatrixFactorization, compute_ground_truth_metrics, evaluate_model
    n, m, d, N = 100_000, 50_000, 64, 1 << 20
    torch.manual_seed(4)
    gt = generate_low_rank_gpu(n, m, d, "cuda", seed=4, force_factored=True)
    ts = sampling.sample_popularity(gt, N, None, "zipf", 1.5)
    assert len(ts) == N and torch.unique(ts.keys).numel() == N
    u, i, j = ts.columns()
    assert float((i < 100).float().mean()) > 0.5                                # popularity bias towards low item ids
    store = sampling.btl_records(gt, ts, scale=5.0, K=4, soft=False, seed=9)
    assert len(store) == 4 * N
    z = store.columns()[3].view(N, 4).mean(1)
    q = torch.sigmoid(5.0 * gt.scale * ((gt.A[u] * gt.B[i]).sum(1) - (gt.A[u] * gt.B[j]).sum(1)))
    assert abs(float((z - q).mean())) < 2e-3 and float((z - q).abs().mean()) < 0.2
    loader = TripletLoader(store, 64)
    gl, ga = compute_ground_truth_metrics(loader, gt, "cuda")
    assert 0.5 < ga <= 1.0 and 0.0 < gl < 0.3
    torch.manual_seed(1)
    model = MatrixFactorization(n, m, d)
    loss_eval, acc = evaluate_model(model, TripletLoader(store, len(store)), "cuda")     # one batch
    fs = model.flat_state(G.DEV)
    gU = torch.zeros_like(fs.U); gV = torch.zeros_like(fs.V); loss = torch.zeros(1, device=G.DEV)
    check(lib.mfcd_triplet_fwd_bwd(ptr(fs.U), ptr(fs.V), ptr(store.rec), None, 0, len(store), d, 1.0 / len(store),
                                   ptr(gU), ptr(gV), ptr(loss), current_stream()), "k1")
    assert abs(loss.item() - loss_eval) < 1e-4 * loss_eval and 0.3 < acc < 0.7


def test_c5_shape_full_metrics_block(G):
    """config 5 width (20 000 items) with the tensor-core statistics engine and Spearman on a 1024-row block:
    a model equal to the ground-truth factors has alpha = 1, Pearson = Spearman = 1."""
    import structure
    from mfcd_b200.store import GroundTruth
    from mfcd_b200.trainer import MatrixFactorization
    n, m, d = 1024, 20_000, 32
    torch.manual_seed(5)
    model = MatrixFactorization(n, m, d)
    fs = model.flat_state(G.DEV)
    X = (fs.U.view(n, d) @ fs.V.view(m, d).T).contiguous()
    out = structure.compute_alpha_and_norm_ratios(model, X)
    assert abs(out[0] - 1) < 1e-4 and out[3] < 1e-3 and abs(out[4] - 1) < 1e-5 and abs(out[6] - 1) < 1e-5
    rec = structure.compute_reconstruction_error(model, X, 1.0)
    col_mean_norm = float(torch.linalg.norm(X.mean(0, keepdim=True).expand_as(X)) / torch.linalg.norm(X))
    assert abs(rec - col_mean_norm) < 1e-3 * max(col_mean_norm, 1e-3)          # only the column means differ
