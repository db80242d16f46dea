"""ctypes binding of libmfcd_b200.so -- the C ABI in include/mfcd_b200.h.

There is no CPU fallback: if the shared library is missing this module raises
at import time, and every compute entry point returns an error without a CUDA
device.  PyTorch is used only for device memory, streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libmfcd_b200.so"


class MfcdError(RuntimeError):
    pass


def library_path() -> str:
    return os.environ.get("MFCD_B200_LIB", os.path.join(_HERE, _LIB_NAME))


class XView(C.Structure):
    """struct mfcd_xview"""
    _fields_ = [("X", C.c_void_p), ("ldx", C.c_int64), ("A", C.c_void_p), ("B", C.c_void_p),
                ("dx", C.c_int32), ("scale", C.c_float)]


class EpochArgs(C.Structure):
    """struct mfcd_epoch_args"""
    _fields_ = [
        ("params", C.c_void_p), ("grads", C.c_void_p), ("state1", C.c_void_p), ("state2", C.c_void_p),
        ("n_users", C.c_int64), ("n_items", C.c_int64),
        ("d", C.c_int32), ("optimizer", C.c_int32), ("mode", C.c_int32), ("flags", C.c_int32),
        ("rec", C.c_void_p), ("perm", C.c_void_p),
        ("n_samples", C.c_int64), ("batch_size", C.c_int64),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("weight_decay", C.c_float), ("momentum", C.c_float),
        ("step0", C.c_int64), ("step_losses", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
        ("item_slot", C.c_void_p), ("hot_items", C.c_void_p), ("n_hot", C.c_int32), ("reserved2", C.c_int32),
    ]


P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
U64 = C.c_uint64
F32 = C.c_float
SZ = C.c_size_t

# name -> argtypes; every function returns int.  Mirrors include/mfcd_b200.h one to one
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "mfcd_device_sm_count": [C.POINTER(C.c_int)],
    "mfcd_profile_k1": [I32],
    "mfcd_profile_k1_read": [C.POINTER(C.c_double), C.POINTER(I64)],
    "mfcd_pack_triplets": [P, P, P, P, I64, P, P],
    "mfcd_unpack_triplets": [P, I64, P, P, P, P, P],
    "mfcd_pack_triplets8": [P, I64, P, P, P],
    "mfcd_unpack_triplets8": [P, I64, P, P],
    "mfcd_host_pack_triplets8": [P, I64, P, I32, P],
    "mfcd_host_pack_isa": [],
    "mfcd_wire_layout": [I64, C.POINTER(I64), C.POINTER(I64), C.POINTER(SZ)],
    "mfcd_pack_wire": [P, I64, P, I64, P, P, SZ, P],
    "mfcd_unpack_wire": [P, I64, P, P],
    "mfcd_gather_triplets": [P, P, I64, P, P],
    "mfcd_epoch_max_batches": [C.POINTER(I32)],
    "mfcd_epoch_positions": [I64, U64, P, P],
    "mfcd_invert_perm": [P, I64, P, P],
    "mfcd_epoch_batches_workspace": [I64, I64, C.POINTER(SZ)],
    "mfcd_epoch_batches": [P, I64, I64, P, U64, P, P, SZ, P],
    "mfcd_triplet_fwd_bwd": [P, P, P, P, I64, I64, I32, F32, P, P, P, P],
    "mfcd_max_hot_items": [I32, C.POINTER(I32)],
    "mfcd_triplet_fwd_bwd_hot": [P, P, P, P, I64, I64, I32, F32, P, P, P, P, P, I32, P],
    "mfcd_triplet_fwd_bwd_ex": [P, P, P, P, I64, I64, I32, F32, P, P, P, P, P, I32, I32, P],
    "mfcd_group_by_user_workspace": [I64, I64, C.POINTER(SZ)],
    "mfcd_group_by_user": [P, I64, I64, P, SZ, P],
    "mfcd_det_workspace_bytes": [I64, I32, C.POINTER(SZ)],
    "mfcd_det_workspace_bytes_nm": [I64, I32, I64, I64, C.POINTER(SZ)],
    "mfcd_triplet_fwd_bwd_det": [P, P, P, P, I64, I64, I32, F32, I64, I64, P, P, P, P, SZ, P],
    "mfcd_triplet_fwd_bwd_det_sort": [P, P, P, P, I64, I64, I32, F32, I64, I64, P, P, P, P, SZ, P],
    "mfcd_triplet_fwd_bwd_det_fixed": [P, P, P, P, I64, I64, I32, F32, I64, I64, P, P, P, P, SZ, P],
    "mfcd_det_fixed_workspace_bytes": [I32, I64, I64, C.POINTER(SZ)],
    "mfcd_adam_update": [P, P, P, P, I64, F32, F32, F32, F32, F32, I64, I32, P],
    "mfcd_sgd_update": [P, P, P, I64, F32, F32, F32, I64, I32, P],
    "mfcd_dp_shard_range": [I64, I32, I32, C.POINTER(I64), C.POINTER(I64)],
    "mfcd_dp_fused_adam": [C.POINTER(U64), C.POINTER(U64), U64, U64, I32, I32, I64, P, P, F32, F32, F32, F32, F32,
                           I64, P],
    "mfcd_dp_fused_adam_sync": [C.POINTER(U64), C.POINTER(U64), C.POINTER(U64), U64, U64, I32, I32, I64, P, P, F32,
                                F32, F32, F32, F32, I64, C.c_uint32, P, P, P, P],
    "mfcd_train_epoch": [C.POINTER(EpochArgs)],
    "mfcd_train_epoch_workspace": [C.POINTER(EpochArgs), C.POINTER(SZ)],
    "mfcd_triplet_eval": [P, P, P, I64, I32, I64, P, P, P, P],
    "mfcd_ground_truth_eval": [C.POINTER(XView), P, I64, I64, P, P, P, P],
    "mfcd_triplet_scores": [P, P, P, P, P, I64, I32, P, P],
    "mfcd_sample_random": [I64, I64, I64, U64, U64, P, P],
    "mfcd_sample_margin": [I64, I64, I64, U64, U64, C.POINTER(XView), F32, P, P],
    "mfcd_sample_popularity": [I64, I64, I64, U64, U64, P, P, P],
    "mfcd_sample_block": [I64, I64, U64, U64, P, I64, P, I64, P, P],
    "mfcd_sample_lists": [I64, I64, I64, U64, U64, P, I32, P, I32, I32, P, P],
    "mfcd_unique_workspace_bytes": [I64, I64, C.POINTER(SZ)],
    "mfcd_unique_accept": [P, I64, P, I64, I64, P, P, P, SZ, P],
    "mfcd_btl_labels": [C.POINTER(XView), P, I64, I64, I32, F32, I32, U64, P, P, P],
    "mfcd_philox_uniforms": [U64, U64, I64, P, P],
    "mfcd_table_col_means": [P, I64, I32, P, P],
    "mfcd_recon_stats": [P, P, I64, I64, I32, C.POINTER(XView), F32, P, P, P, P],
    "mfcd_recon_stats_tc_workspace_bytes": [I64, I64, I32, C.POINTER(SZ)],
    "mfcd_recon_stats_tc": [P, P, I64, I64, I32, C.POINTER(XView), F32, P, P, P, P, P, SZ, P],
    "mfcd_reconstruct_rows": [P, P, I64, I64, I64, I32, P, P],
    "mfcd_xview_rows": [C.POINTER(XView), I64, I64, I64, P, P],
    "mfcd_rank_workspace_bytes": [I64, I64, C.POINTER(SZ)],
    "mfcd_row_ranks": [P, I64, I64, P, P, SZ, P],
    "mfcd_row_pearson": [P, P, I64, I64, P, P],
}


def _load():
    path = library_path()
    if not os.path.exists(path):
        raise MfcdError(
            f"{path} not found. The CUDA extension is the product path and there is no fallback: "
            f"build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C {os.path.join(_HERE, 'csrc')}`).")
    try:
        handle = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise MfcdError(f"could not load {path}: {e}") from e
    handle.mfcd_abi_version.restype = C.c_int
    handle.mfcd_abi_version.argtypes = []
    handle.mfcd_last_error.restype = C.c_char_p
    handle.mfcd_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name)       # AttributeError here = header / library drift
        fn.restype = C.c_int
        fn.argtypes = argtypes
    return handle


lib = _load()
ABI_VERSION = lib.mfcd_abi_version()
FLAG_USER_GROUPED = 1      # MFCD_FLAG_USER_GROUPED
FLAG_WIRE_RLE = 2          # MFCD_FLAG_WIRE_RLE


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.mfcd_last_error().decode("utf-8", "replace")
        raise MfcdError(f"{what or 'mfcd call'} failed (rc={rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def wait_for_stream(dev=None):
    """Block until the current stream of `dev` is idle.  Stream.synchronize releases the GIL; a thread that waits for
    the GPU inside Tensor.tolist() / .item() / .cpu() instead kept every other Python thread of a concurrent sweep
    (structure.parameter_scan(concurrency=k)) from running for the length of its GPU work."""
    import torch
    with torch.cuda.device(dev):
        torch.cuda.current_stream().synchronize()
