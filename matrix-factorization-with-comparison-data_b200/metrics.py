"""Reconstruction / alignment / correlation metrics (kernels K5 and K6).

Mirrors ``compute_reconstruction_error`` (structure.py:925-955) and
``compute_alpha_and_norm_ratios`` (structure.py:958-1082).  The reference
materialises UV^T, copies it to numpy and loops over rows in python; here one
streaming pass (mfcd_recon_stats) leaves 6 fp64 sums per row, from which every
scalar and per-row list is finished on the host in float64, and Spearman comes
from GPU row ranks (mfcd_row_ranks + mfcd_row_pearson) over row blocks.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from ._lib import lib, check, ptr, current_stream, wait_for_stream, MfcdError
from .store import GroundTruth, compute_device


def _row_range(n, world_size):
    """(rank, world, lo, hi): the rows of X this rank evaluates (all of them on one GPU)"""
    from .trainer import dist_world
    from .dist import split_even
    rank, world = dist_world(world_size)
    lo, hi = split_even(n, world, rank) if world > 1 else (0, n)
    return rank, world, lo, hi


def _combine_rows(t, world):
    """rows were computed by their owners into a zero-filled tensor: sum over the ranks = the full tensor"""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def _row_stats(model, gt: GroundTruth, s: float, engine=None, world_size=1):
    """(n x 8) fp64 row statistics of mfcd_recon_stats as a numpy array, and the model's flat state."""
    stats, fs = _row_stats_device(model, gt, s, engine=engine, world_size=world_size)
    wait_for_stream(stats.device)
    return stats.cpu().numpy(), fs


def _row_stats_device(model, gt: GroundTruth, s: float, engine=None, world_size=1):
    """(n x 8) fp64 row statistics of mfcd_recon_stats (device tensor).  Data parallel: each rank streams only
    its rows of X (K5 shards rows of X), the per-row results are combined with one all-reduce."""
    fs = model.flat_state(gt.device)
    dev = fs.params.device
    n, m, d = fs.n, fs.m, fs.d
    assert gt.shape == (n, m), f"X is {gt.shape}, model is {(n, m)}"
    rank, world, lo, hi = _row_range(n, world_size)
    nl = hi - lo
    ubar = torch.empty(d, dtype=torch.float32, device=dev)
    vbar = torch.empty(d, dtype=torch.float32, device=dev)
    stats = torch.zeros((n, 8), dtype=torch.float64, device=dev) if world > 1 else \
        torch.empty((n, 8), dtype=torch.float64, device=dev)
    gl = gt.row_slice(lo, hi) if world > 1 else gt
    xv = gl.xview()
    Ul = fs.U[lo * d:]
    sl = stats[lo:]
    with torch.cuda.device(dev):
        st = current_stream()
        check(lib.mfcd_table_col_means(ptr(fs.U), n, d, ptr(ubar), st), "mfcd_table_col_means(U)")
        check(lib.mfcd_table_col_means(ptr(fs.V), m, d, ptr(vbar), st), "mfcd_table_col_means(V)")
        engine = os.environ.get("MFCD_K5", "auto") if engine is None else engine
        done = nl == 0
        if not done and engine in ("auto", "tc"):
            # tensor-core path (tcgen05 / TMEM) when the shape is eligible; worth it once K = d is big enough
            # for the fp32 FMA pipe to be the limiter of the SIMT engine
            if engine == "tc" or d >= 16:
                flag = torch.zeros(1, dtype=torch.int32, device=dev)
                need = C.c_size_t(0)
                check(lib.mfcd_recon_stats_tc_workspace_bytes(nl, m, d, C.byref(need)), "mfcd_recon_stats_tc_workspace_bytes")
                ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
                rc = lib.mfcd_recon_stats_tc(ptr(Ul), ptr(fs.V), nl, m, d, C.byref(xv), float(s), ptr(ubar),
                                             ptr(vbar), ptr(sl), ptr(flag), ptr(ws), need.value, st)
                if rc == 0:
                    wait_for_stream(dev)
                    if int(flag.item()) == 0:
                        done = True
                    elif engine == "tc":
                        raise MfcdError("mfcd_recon_stats_tc: tensor-core pipeline timed out")
                    else:       # do not lose a finished training run to a metrics kernel: recompute exactly
                        import warnings
                        warnings.warn("mfcd_recon_stats_tc reported a pipeline time-out; recomputing with the "
                                      "SIMT engine (mfcd_recon_stats)")
                elif rc != -3 or engine == "tc":          # -3 = MFCD_ERR_UNSUPPORTED -> SIMT engine
                    check(rc, "mfcd_recon_stats_tc")
        if not done:
            check(lib.mfcd_recon_stats(ptr(Ul), ptr(fs.V), nl, m, d, C.byref(xv), float(s), ptr(ubar), ptr(vbar),
                                       ptr(sl), st), "mfcd_recon_stats")
        _combine_rows(stats, world)
    return stats, fs


def compute_reconstruction_error(model, X, s, *, world_size=1):
    """|| (UV^T - column means) - sX ||_F / || sX ||_F   (structure.py:939-955)."""
    gt = GroundTruth.wrap(X)
    st, _ = _row_stats(model, gt, s, world_size=world_size)
    num = math.sqrt(float(st[:, 5].sum()))
    den = abs(float(s)) * math.sqrt(float(st[:, 1].sum()))
    return num / den


def _centered_sums(st, m):
    sx, sxx, sw, sww, sxw = st[:, 0], st[:, 1], st[:, 2], st[:, 3], st[:, 4]
    cxx = np.maximum(sxx - sx * sx / m, 0.0)     # sum (x - xbar)^2
    cww = np.maximum(sww - sw * sw / m, 0.0)     # sum (w - wbar)^2   (w already centred by <U_r, vbar>)
    cxw = sxw - sx * sw / m                      # sum (x - xbar)(w - wbar)
    return cxx, cww, cxw


def _singular_values_lowrank(L, R):
    """Singular values of L @ R.T for tall-skinny L (n x k), R (m x k) via two QRs."""
    _, rl = torch.linalg.qr(L.double(), mode="r")
    _, rr = torch.linalg.qr(R.double(), mode="r")
    return torch.linalg.svdvals(rl @ rr.T)


def _svd_error(fs, gt: GroundTruth, alpha, norm_X):
    """|| alpha S(W_c) - S(X_c) ||_2 / (|| S(X_c) ||_2 + 1e-8)   (structure.py:1013-1017).
    S(W_c): W_c = U (V - vbar)^T has rank <= d, so its singular values come from a
    d x d core (library QR/SVD on tiny matrices); the remaining min(n,m)-d are 0.
    S(X_c): low-rank X by the same trick, dense X by torch.linalg.svdvals (library)."""
    n, m, d = fs.n, fs.m, fs.d
    U = fs.U.view(n, d)
    V = fs.V.view(m, d)
    k = min(n, m)
    s2 = _singular_values_lowrank(U, V - V.mean(dim=0, keepdim=True))
    S2 = torch.zeros(k, dtype=torch.float64, device=U.device)
    S2[: min(k, s2.numel())] = s2[:k]
    if gt.X is None or gt.factors is not None:
        # low-rank ground truth (factored, or a dense X whose generator left its factors): X - rowmean(X) =
        # scale * A (B - colmean(B))^T, singular values from a dx x dx core (reference: structure.py:1011-1017
        # runs a full dense SVD here)
        A, B, scale = (gt.A, gt.B, gt.scale) if gt.X is None else gt.factors
        s1 = _singular_values_lowrank(A * scale, B - B.mean(dim=0, keepdim=True))
        S1 = torch.zeros(k, dtype=torch.float64, device=U.device)
        S1[: min(k, s1.numel())] = s1[:k]
    else:
        Xc = gt.X - gt.X.mean(dim=1, keepdim=True)
        S1 = torch.linalg.svdvals(Xc).double()[:k]
    diff = alpha * S2 - S1
    return float(torch.linalg.norm(diff) / (torch.linalg.norm(S1) + 1e-8))


def row_spearman(model, gt: GroundTruth, rows_mask=None, max_bytes=2 << 30, world_size=1):
    """Spearman rho of every row of X against the same row of UV^T (float64
    numpy array of length n; NaN where a row is constant).  Data parallel: rows are split over the ranks."""
    fs = model.flat_state(gt.device)
    dev = fs.params.device
    n, m, d = fs.n, fs.m, fs.d
    rank, world, lo, hi = _row_range(n, world_size)
    rho = torch.zeros(n, dtype=torch.float64, device=dev)
    per_row = 4 * m
    one = C.c_size_t(0)
    check(lib.mfcd_rank_workspace_bytes(1, m, C.byref(one)), "mfcd_rank_workspace_bytes")
    chunk = int(max(1, min(n, max_bytes // (4 * per_row + max(one.value, 1)))))
    ws_bytes = C.c_size_t(0)
    check(lib.mfcd_rank_workspace_bytes(chunk, m, C.byref(ws_bytes)), "mfcd_rank_workspace_bytes")
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    wrows = torch.empty((chunk, m), dtype=torch.float32, device=dev)
    rx = torch.empty((chunk, m), dtype=torch.float32, device=dev)
    rw = torch.empty((chunk, m), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = current_stream()
        for r0 in range(lo, hi, chunk):
            nr = min(chunk, hi - r0)
            xrows = gt.rows(r0, nr).contiguous()
            check(lib.mfcd_reconstruct_rows(ptr(fs.U), ptr(fs.V), r0, nr, m, d, ptr(wrows), st),
                  "mfcd_reconstruct_rows")
            check(lib.mfcd_row_ranks(ptr(xrows), nr, m, ptr(rx), ptr(ws), ws.numel(), st), "mfcd_row_ranks(X)")
            check(lib.mfcd_row_ranks(ptr(wrows), nr, m, ptr(rw), ptr(ws), ws.numel(), st), "mfcd_row_ranks(W)")
            check(lib.mfcd_row_pearson(ptr(rx), ptr(rw), nr, m, ptr(rho[r0:]), st), "mfcd_row_pearson")
        _combine_rows(rho, world)
    wait_for_stream(rho.device)
    return rho.cpu().numpy()


def compute_alpha_and_norm_ratios(model, X_init, *, world_size=1):
    """The reference's 14-tuple (structure.py:958-1082), same order and types:
    alpha, norm_X, norm_ratio, reconstruction_error_scaled, pearson_mean, pearson_std,
    spearman_mean, spearman_std, svd_error_scaled, slopes, correlations,
    spearman_scores, reconstruction_error_scaled_per_row, alpha_per_row."""
    gt = GroundTruth.wrap(X_init)
    st, fs = _row_stats(model, gt, 1.0, world_size=world_size)
    n, m = gt.shape
    cxx, cww, cxw = _centered_sums(st, float(m))

    dot = float(cxw.sum())
    norm_w2 = float(cww.sum())
    norm_w = math.sqrt(norm_w2)
    norm_X = math.sqrt(float(cxx.sum()))
    alpha = dot / (norm_w2 + 1e-8)
    norm_ratio = norm_w / (norm_X + 1e-8)
    rec_scaled = math.sqrt(max(alpha * alpha * norm_w2 - 2.0 * alpha * dot + norm_X * norm_X, 0.0)) / (norm_X + 1e-8)

    std_x = np.sqrt(cxx / m)
    std_w = np.sqrt(cww / m)
    ok = (std_x > 1e-8) & (std_w > 1e-8)                       # structure.py:1006
    with np.errstate(divide="ignore", invalid="ignore"):
        corr_all = cxw / np.sqrt(cxx * cww)
    correlations = [float(c) for c in corr_all[ok]]
    pearson_mean = float(np.mean(correlations)) if correlations else 0.0

    try:
        svd_error_scaled = _svd_error(fs, gt, alpha, norm_X)
    except Exception:                                          # the reference's bare except (:1018-1020)
        pearson_mean = 0.0
        svd_error_scaled = 1.0

    spearman_scores = []
    if ok.any():
        rho = row_spearman(model, gt, world_size=world_size)
        spearman_scores = [float(r) for r in rho[ok] if not math.isnan(r)]
    spearman_mean = float(np.mean(spearman_scores)) if spearman_scores else 0.0
    pearson_std = float(np.std(correlations)) if correlations else 0.0
    spearman_std = float(np.std(spearman_scores)) if spearman_scores else 0.0

    ok_slope = (cxx > 1e-8) & (std_w > 1e-8)                   # structure.py:1042-1043
    with np.errstate(divide="ignore", invalid="ignore"):
        slopes = [float(v) for v in (cxw / cxx)[ok_slope]]
        alpha_rows = np.where(cww > 1e-8, cxw / np.where(cww > 1e-8, cww, 1.0), 0.0)   # :1057-1058
    alpha_per_row = [float(a) for a in alpha_rows]
    rec_row_sq = float((alpha_rows * alpha_rows * cww - 2.0 * alpha_rows * cxw + cxx).sum())
    rec_row = math.sqrt(max(rec_row_sq, 0.0)) / (norm_X + 1e-8)

    return (alpha, norm_X, norm_ratio, rec_scaled, pearson_mean, pearson_std, spearman_mean, spearman_std,
            svd_error_scaled, slopes, correlations, spearman_scores, rec_row, alpha_per_row)


def sampled_rows(model, X, row_indices):
    """Rows of UV^T and of X for visual inspection (structure.py:388-392)."""
    gt = GroundTruth.wrap(X)
    fs = model.flat_state(gt.device)
    dev = fs.params.device
    n, m, d = fs.n, fs.m, fs.d
    xs, ws = [], []
    with torch.cuda.device(dev):
        for r in row_indices:
            r = int(r)
            out = torch.empty((1, m), dtype=torch.float32, device=dev)
            check(lib.mfcd_reconstruct_rows(ptr(fs.U), ptr(fs.V), r, 1, m, d, ptr(out), current_stream()),
                  "mfcd_reconstruct_rows")
            ws.append(out[0].cpu().numpy())
            xs.append(gt.rows(r, 1)[0].cpu().numpy())
    return np.stack(xs), np.stack(ws)
