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
