"""Helpers shared by the GPU tests: thin wrappers that call the C ABI directly."""
import ctypes as C

import numpy as np
import torch

import mfcd_b200
from mfcd_b200 import _lib
from mfcd_b200._lib import lib, check, ptr, current_stream
from mfcd_b200.store import TripletStore

DEV = torch.device("cuda", 0)


def dev_f32(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(DEV)


def store_from(u, i, j, z):
    return TripletStore.from_columns(torch.from_numpy(np.asarray(u, np.int64)), torch.from_numpy(np.asarray(i, np.int64)),
                                     torch.from_numpy(np.asarray(j, np.int64)), torch.from_numpy(np.asarray(z, np.float64)),
                                     device=DEV)


def fwd_bwd(U, V, store, start=0, B=None, mode="atomic", perm=None, inv_batch=None, gU=None, gV=None):
    """-> (loss float, gU ndarray, gV ndarray) through mfcd_triplet_fwd_bwd{,_det}."""
    n, d = U.shape
    m = V.shape[0]
    B = len(store) - start if B is None else B
    inv = (1.0 / max(B, 1)) if inv_batch is None else inv_batch
    Ud, Vd = dev_f32(U), dev_f32(V)
    gUd = torch.zeros_like(Ud) if gU is None else dev_f32(gU)
    gVd = torch.zeros_like(Vd) if gV is None else dev_f32(gV)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)
    permd = None if perm is None else torch.as_tensor(perm).to(DEV, torch.int32)
    if mode == "atomic":
        check(lib.mfcd_triplet_fwd_bwd(ptr(Ud), ptr(Vd), ptr(store.rec), ptr(permd), start, B, d, inv, ptr(gUd),
                                       ptr(gVd), ptr(loss), current_stream()), "fwd_bwd")
    elif mode == "deterministic_sort":      # the sort + segmented-reduction engine under its own entry point
        need = C.c_size_t(0)
        check(lib.mfcd_det_workspace_bytes(B, d, C.byref(need)), "ws")
        ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=DEV)
        check(lib.mfcd_triplet_fwd_bwd_det_sort(ptr(Ud), ptr(Vd), ptr(store.rec), ptr(permd), start, B, d, inv, n, m,
                                                ptr(gUd), ptr(gVd), ptr(loss), ptr(ws), need.value, current_stream()),
              "fwd_bwd_det_sort")
    elif mode == "deterministic_fixed":     # the fixed-point integer-atomic engine under its own entry point
        need = C.c_size_t(0)
        check(lib.mfcd_det_fixed_workspace_bytes(d, n, m, C.byref(need)), "ws")
        ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=DEV)
        check(lib.mfcd_triplet_fwd_bwd_det_fixed(ptr(Ud), ptr(Vd), ptr(store.rec), ptr(permd), start, B, d, inv, n, m,
                                                 ptr(gUd), ptr(gVd), ptr(loss), ptr(ws), need.value, current_stream()),
              "fwd_bwd_det_fixed")
    else:                                   # default engine for this batch size
        need = C.c_size_t(0)
        check(lib.mfcd_det_workspace_bytes_nm(B, d, n, m, C.byref(need)), "ws")
        ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=DEV)
        check(lib.mfcd_triplet_fwd_bwd_det(ptr(Ud), ptr(Vd), ptr(store.rec), ptr(permd), start, B, d, inv, n, m,
                                           ptr(gUd), ptr(gVd), ptr(loss), ptr(ws), need.value, current_stream()),
              "fwd_bwd_det")
    torch.cuda.synchronize()
    return float(loss.item()), gUd.cpu().numpy(), gVd.cpu().numpy()


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
