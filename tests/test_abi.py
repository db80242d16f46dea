"""The C-ABI library loads without a GPU and exports exactly what include/mfcd_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mfcd_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(mfcd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        nargs = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[name] = nargs
    return out


def test_header_declares_the_hot_path():
    fns = declared_functions()
    for required in ("mfcd_triplet_fwd_bwd", "mfcd_triplet_fwd_bwd_det", "mfcd_adam_update", "mfcd_sgd_update",
                     "mfcd_train_epoch", "mfcd_triplet_eval", "mfcd_ground_truth_eval", "mfcd_recon_stats",
                     "mfcd_row_ranks", "mfcd_sample_random", "mfcd_sample_margin", "mfcd_sample_popularity",
                     "mfcd_sample_block", "mfcd_unique_accept", "mfcd_btl_labels"):
        assert required in fns


def test_library_exports_every_declared_symbol():
    import mfcd_b200
    from mfcd_b200 import _lib
    handle = C.CDLL(_lib.library_path())
    for name in declared_functions():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert _lib.ABI_VERSION == 1


def test_binding_matches_header_arity():
    from mfcd_b200 import _lib
    fns = declared_functions()
    bound = dict(_lib.SIGNATURES)
    bound["mfcd_abi_version"] = []
    bound["mfcd_last_error"] = []
    assert set(bound) == set(fns), set(bound) ^ set(fns)
    for name, argtypes in bound.items():
        assert len(argtypes) == fns[name], (name, len(argtypes), fns[name])


def test_struct_layouts_match_the_header():
    from mfcd_b200 import _lib
    assert C.sizeof(_lib.XView) == 40           # ptr, i64, ptr, ptr, i32, f32
    assert _lib.EpochArgs.rec.offset == 64 and _lib.EpochArgs.step0.offset == 120
    assert C.sizeof(_lib.EpochArgs) == 184 and _lib.EpochArgs.item_slot.offset == 160


def test_argument_errors_are_reported_without_a_gpu():
    from mfcd_b200 import _lib
    n = C.c_size_t(0)
    assert _lib.lib.mfcd_det_workspace_bytes(-1, 4, C.byref(n)) == -1
    assert b"bad argument" in _lib.lib.mfcd_last_error()
    assert _lib.lib.mfcd_det_workspace_bytes(64, 4, C.byref(n)) == 0 and n.value == 0     # single-CTA path
    assert _lib.lib.mfcd_adam_update(None, None, None, None, -5, 0.1, 0.9, 0.999, 1e-8, 0.0, 1, 1, None) == -1
    with pytest.raises(_lib.MfcdError):
        _lib.check(-1, "probe")


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import importlib
    from mfcd_b200 import _lib
    monkeypatch.setenv("MFCD_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MfcdError):
        _lib._load()


def _epoch_ws(n, m, d, batch, N, mode=1, opt=0):
    from mfcd_b200 import _lib
    a = _lib.EpochArgs()
    a.n_users, a.n_items, a.d = n, m, d
    a.mode, a.optimizer = mode, opt
    a.batch_size, a.n_samples = batch, N
    a.beta1, a.beta2 = 0.9, 0.999
    out = C.c_size_t(0)
    assert _lib.lib.mfcd_train_epoch_workspace(C.byref(a), C.byref(out)) == 0
    return out.value


def test_epoch_workspace_says_when_the_persistent_kernel_applies():
    """Host-side plan of mfcd_train_epoch (no GPU needed): the reference's regime -- deterministic, Adam, batch <= 256,
    small tables -- asks for the second parameter buffer + the per-step bias table; everything else asks for
    nothing (atomic mode, SGD, huge tables) or for the sort workspace of the large deterministic path."""
    n, m, d, N = 1000, 1000, 10, 600_000
    steps = (N + 63) // 64
    need = _epoch_ws(n, m, d, 64, N)
    assert 4 * (n + m) * d + 8 * steps <= need <= 4 * (n + m) * d + 8 * steps + 3 * 256
    assert _epoch_ws(100, 100, 2, 64, 400) >= 4 * 400 + 8 * 7          # config 1: 7 steps per epoch
    assert _epoch_ws(n, m, d, 64, N, mode=0) == 0                      # atomic mode: per-step launches
    assert _epoch_ws(n, m, d, 64, N, opt=1) == 0                       # SGD: per-step launches
    assert _epoch_ws(100_000, 50_000, 64, 64, 1000) == 0               # 9.6 M elements do not fit the register slices
    big = _epoch_ws(n, m, d, 512, N)                                   # batch > 256: sort + segmented reduction
    assert big > 0 and big != need
    assert _epoch_ws(101, 100, 3, 64, 400) > 0                         # odd d: scalar rows, still eligible
