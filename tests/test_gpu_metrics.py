"""GPU parity tests for the evaluation metrics (K5 recon stats, K6 row ranks):
1e-3 on final metrics (BASELINE.json north_star), against goldens from the reference."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mfcd_oracle as O

pytestmark = pytest.mark.gpu

TRAIN_FIXTURES = ["train_c1.npz", "train_d10_k3.npz", "train_soft_d4.npz", "train_d64.npz"]


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


def _model(g, which="1"):
    from mfcd_b200.trainer import MatrixFactorization
    model = MatrixFactorization(int(g["n"]), int(g["m"]), int(g["d"]))
    with torch.no_grad():
        model.U.copy_(torch.from_numpy(g["U" + which])); model.V.copy_(torch.from_numpy(g["V" + which]))
    return model


def _check_tuple(out, g, tol=1e-3):
    scal = [out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], out[8], out[12]]
    for k, (a, b) in enumerate(zip(scal, g["alpha_scalars"])):
        assert abs(a - b) <= tol * max(abs(b), 1e-3), (k, a, b)
    assert len(out[9]) == len(g["slopes"]) and np.abs(np.array(out[9]) - g["slopes"]).max() <= tol * max(1e-3, np.abs(g["slopes"]).max())
    assert len(out[10]) == len(g["correlations"]) and np.abs(np.array(out[10]) - g["correlations"]).max() < tol
    assert len(out[11]) == len(g["spearman_scores"]) and np.abs(np.array(out[11]) - g["spearman_scores"]).max() < tol
    assert len(out[13]) == len(g["alpha_per_row"])
    assert np.abs(np.array(out[13]) - g["alpha_per_row"]).max() <= tol * max(1e-3, np.abs(g["alpha_per_row"]).max())
    assert all(isinstance(v, float) for v in scal) and isinstance(out[9], list) and isinstance(out[13], list)


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_metrics_match_reference(G, name):
    import structure
    g = load_golden(name)
    model = _model(g)
    X = torch.from_numpy(g["X"])
    rec = structure.compute_reconstruction_error(model, X, float(g["s"]))
    assert abs(rec - g["rec_err"]) < 1e-3 * g["rec_err"]
    _check_tuple(structure.compute_alpha_and_norm_ratios(model, X), g)


def test_factored_ground_truth_equals_dense(G):
    """X given as low-rank factors (never materialised) gives the same metrics as the dense matrix."""
    import structure
    from mfcd_b200.store import GroundTruth
    g = load_golden("train_d10_k3.npz")
    model = _model(g)
    rng = np.random.default_rng(3)
    n, m = int(g["n"]), int(g["m"])
    A = rng.standard_normal((n, 5)).astype(np.float32); B = rng.standard_normal((m, 5)).astype(np.float32)
    gt_f = GroundTruth(A=torch.from_numpy(A), B=torch.from_numpy(B), scale=0.37)
    Xd = gt_f.dense().cpu()
    assert np.abs(Xd.numpy() - 0.37 * A @ B.T).max() < 1e-5
    a = structure.compute_alpha_and_norm_ratios(model, Xd)
    b = structure.compute_alpha_and_norm_ratios(model, gt_f)
    for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 12):
        assert abs(a[k] - b[k]) <= 1e-4 * max(abs(a[k]), 1e-3), k
    ra = structure.compute_reconstruction_error(model, Xd, 2.0)
    rb = structure.compute_reconstruction_error(model, gt_f, 2.0)
    assert abs(ra - rb) < 1e-5 * ra
    ref = O.reconstruction_error(g["U1"], g["V1"], Xd.numpy(), 2.0)
    assert abs(ra - ref) < 1e-4 * ref


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 2), (65, 63, 3), (130, 257, 33), (64, 64, 64), (200, 70, 130)])
def test_recon_stats_ragged_shapes_against_oracle(G, shape):
    import structure
    from mfcd_b200.trainer import MatrixFactorization
    n, m, d = shape
    rng = np.random.default_rng(sum(shape))
    model = MatrixFactorization(n, m, d)
    X = torch.from_numpy(rng.standard_normal((n, m)).astype(np.float32))
    U, V = model.U.detach().numpy().copy(), model.V.detach().numpy().copy()
    if m > 1:
        rec = structure.compute_reconstruction_error(model, X, 0.7)
        ref = O.reconstruction_error(U, V, X.numpy(), 0.7)
        assert abs(rec - ref) < 1e-4 * ref
    if min(n, m) >= 3:
        out = structure.compute_alpha_and_norm_ratios(model, X)
        ref = O.alpha_and_norm_ratios(U, V, X.numpy())
        for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 12):
            assert abs(out[k] - ref[k]) <= 1e-3 * max(abs(ref[k]), 1e-3), (k, out[k], ref[k])
        assert np.abs(np.array(out[11]) - np.array(ref[11])).max() < 1e-3


def test_row_ranks_with_ties_and_pearson(G):
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(0)
    rows, m = 37, 501
    vals = rng.integers(0, 40, (rows, m)).astype(np.float32)        # heavy ties
    vals[0] = 3.0                                                    # constant row
    vals[1] = np.arange(m)                                           # strictly increasing
    vals[2] = -np.arange(m)
    vd = torch.from_numpy(vals).to(G.DEV)
    ranks = torch.empty_like(vd)
    need = C.c_size_t(0)
    check(lib.mfcd_rank_workspace_bytes(rows, m, C.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=G.DEV)
    check(lib.mfcd_row_ranks(ptr(vd), rows, m, ptr(ranks), ptr(ws), need.value, current_stream()), "ranks")
    r = ranks.cpu().numpy()
    for k in range(rows):
        assert np.array_equal(r[k], O.average_ranks(vals[k]).astype(np.float32)), k
    other = torch.from_numpy(rng.standard_normal((rows, m)).astype(np.float32)).to(G.DEV)
    r2 = torch.empty_like(other)
    check(lib.mfcd_row_ranks(ptr(other), rows, m, ptr(r2), ptr(ws), need.value, current_stream()), "ranks")
    rho = torch.empty(rows, dtype=torch.float64, device=G.DEV)
    check(lib.mfcd_row_pearson(ptr(ranks), ptr(r2), rows, m, ptr(rho), current_stream()), "pearson")
    rho = rho.cpu().numpy()
    assert math.isnan(rho[0])                                        # constant row -> NaN like scipy
    for k in range(1, rows):
        assert abs(rho[k] - O.pearson(O.average_ranks(vals[k]), O.average_ranks(other[k].cpu().numpy()))) < 1e-9
    with pytest.raises(Exception):
        check(lib.mfcd_row_ranks(ptr(vd), rows, m, ptr(ranks), ptr(ws), 16, current_stream()), "ranks")


def test_degenerate_rows_are_skipped_like_the_reference(G):
    """rows with zero variance are dropped from the Pearson / Spearman lists, alpha_i = 0 where <w,w> <= 1e-8."""
    import structure
    from mfcd_b200.trainer import MatrixFactorization
    n, m, d = 12, 40, 3
    model = MatrixFactorization(n, m, d)
    with torch.no_grad():
        model.U[3].zero_()                       # row 3 of UV^T is identically 0
    rng = np.random.default_rng(2)
    X = rng.standard_normal((n, m)).astype(np.float32)
    X[5] = 1.25                                  # constant ground-truth row
    out = structure.compute_alpha_and_norm_ratios(model, torch.from_numpy(X))
    ref = O.alpha_and_norm_ratios(model.U.detach().cpu().numpy(), model.V.detach().cpu().numpy(), X)
    assert len(out[10]) == len(ref[10]) == n - 2 and len(out[11]) == len(ref[11])
    assert len(out[9]) == len(ref[9]) and out[13][3] == 0.0 and len(out[13]) == n
    for k in (0, 1, 2, 3, 4, 5, 6, 7, 12):
        assert abs(out[k] - ref[k]) <= 1e-3 * max(abs(ref[k]), 1e-3), k


def test_c5_shape_stats_properties(G):
    """BASELINE config 5 width (m = 20000, d = 128) on a row block: W == X = U V^T exactly representable =>
    alpha = 1, per-row Pearson = 1, reconstruction_error_scaled ~ 0."""
    import structure
    from mfcd_b200.trainer import MatrixFactorization
    n, m, d = 512, 20000, 128
    model = MatrixFactorization(n, m, d)
    fs = model.flat_state(G.DEV)
    X = (fs.U.view(n, d) @ fs.V.view(m, d).T).contiguous()
    out = structure.compute_alpha_and_norm_ratios(model, X)
    assert abs(out[0] - 1) < 1e-4 and abs(out[2] - 1) < 1e-4 and out[3] < 1e-3
    assert abs(out[4] - 1) < 1e-5 and abs(out[6] - 1) < 1e-5 and out[8] < 1e-3


@pytest.mark.parametrize("shape", [(128, 64, 8), (300, 260, 32), (1000, 1024, 64), (257, 132, 10), (130, 256, 64),
                                   (64, 4, 3), (513, 1028, 48), (5, 8, 2), (2000, 36, 17), (300, 260, 128),
                                   (1000, 1024, 128), (257, 132, 100), (130, 256, 65), (700, 2052, 96)])
def test_tensor_core_recon_stats_matches_simt(G, shape):
    """tcgen05/TMEM engine (tf32 hi/lo split, 3 MMAs) == fp32 SIMT engine on the six per-row sums."""
    from mfcd_b200 import metrics
    from mfcd_b200.store import GroundTruth
    from mfcd_b200.trainer import MatrixFactorization
    n, m, d = shape
    rng = np.random.default_rng(n + m + d)
    torch.manual_seed(n)
    model = MatrixFactorization(n, m, d)
    gt = GroundTruth(X=torch.from_numpy(rng.standard_normal((n, m)).astype(np.float32)))
    simt, _ = metrics._row_stats(model, gt, 0.7, engine="simt")
    tcst, _ = metrics._row_stats(model, gt, 0.7, engine="tc")
    scale = np.abs(simt[:, :6]).max(axis=0) + 1e-30
    # column 2 is sum(w - rowmean(w)): pure rounding residue around 0, so it is measured against the row's
    # magnitude sqrt(m * sum (w - a)^2) instead of against itself
    scale[2] = np.sqrt(m * simt[:, 3]).max() + 1e-30
    err = np.abs(tcst[:, :6] - simt[:, :6]).max(axis=0) / scale
    assert (err < 2e-5).all(), err
    assert np.abs(tcst[:, 6] - simt[:, 6]).max() < 1e-6


def test_tensor_core_engine_rejects_ineligible_shapes(G):
    from mfcd_b200 import metrics
    from mfcd_b200._lib import MfcdError
    from mfcd_b200.store import GroundTruth
    from mfcd_b200.trainer import MatrixFactorization
    model = MatrixFactorization(40, 30, 128)
    gt = GroundTruth(X=torch.zeros(40, 30))
    with pytest.raises(MfcdError):
        metrics._row_stats(model, gt, 1.0, engine="tc")       # rows of X not 16-byte aligned
    metrics._row_stats(model, gt, 1.0, engine="auto")         # falls back to the SIMT engine
    model = MatrixFactorization(40, 32, 132)
    gt = GroundTruth(X=torch.zeros(40, 32))
    with pytest.raises(MfcdError):
        metrics._row_stats(model, gt, 1.0, engine="tc")       # d > 128
    metrics._row_stats(model, gt, 1.0, engine="auto")


@pytest.mark.parametrize("shape", [(1000, 1024, 128), (1000, 1024, 64), (515, 1284, 32)])
def test_tensor_core_recon_stats_against_the_float64_oracle(G, shape):
    """The tcgen05 engine on multi-tile shapes (A operand in tensor memory, K up to 128) against float64 numpy:
    the six per-row sums directly, and through them the reference's metrics (structure.py:939-955, :980-996)."""
    import structure
    from mfcd_b200 import metrics
    from mfcd_b200.store import GroundTruth
    from mfcd_b200.trainer import MatrixFactorization
    n, m, d = shape
    rng = np.random.default_rng(n * 3 + d)
    torch.manual_seed(d)
    model = MatrixFactorization(n, m, d)
    U, V = model.U.detach().numpy().astype(np.float64), model.V.detach().numpy().astype(np.float64)
    Xf = rng.standard_normal((n, m)).astype(np.float32)
    X = Xf.astype(np.float64)
    s = 0.7
    got, _ = metrics._row_stats(model, GroundTruth(X=torch.from_numpy(Xf)), s, engine="tc")
    W = U @ V.T
    a = U @ V.mean(axis=0)
    b = V @ U.mean(axis=0)
    wa = W - a[:, None]
    want = np.stack([X.sum(1), (X * X).sum(1), wa.sum(1), (wa * wa).sum(1), (X * wa).sum(1),
                     ((W - b[None, :] - s * X) ** 2).sum(1)], 1)
    scale = np.abs(want).max(axis=0)
    scale[2] = np.sqrt(m * want[:, 3]).max()
    err = np.abs(got[:, :6] - want).max(axis=0) / scale
    assert (err < 2e-5).all(), err
    assert np.abs(got[:, 6] - a).max() < 1e-5
    rec = structure.compute_reconstruction_error(model, torch.from_numpy(Xf), s)
    ref = O.reconstruction_error(U.astype(np.float32), V.astype(np.float32), Xf, s)
    assert abs(rec - ref) < 1e-4 * ref
