#!/usr/bin/env python
"""Repeated K5 launches at one shape (separate process per configuration: a launch failure is sticky)."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mfcd_b200._lib import lib, check, ptr, current_stream
from mfcd_b200.store import GroundTruth
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192); ap.add_argument("--m", type=int, default=20480)
ap.add_argument("--d", type=int, default=16); ap.add_argument("--launches", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
U = torch.randn(a.n, a.d, device=dev) / a.d ** 0.5; V = torch.randn(a.m, a.d, device=dev) / a.d ** 0.5
X = torch.randn(a.n, a.m, device=dev)
xv = GroundTruth(X=X).xview()
ubar, vbar = U.mean(0).contiguous(), V.mean(0).contiguous()
stats = torch.empty((a.n, 8), dtype=torch.float64, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
need = C.c_size_t(0)
check(lib.mfcd_recon_stats_tc_workspace_bytes(a.n, a.m, a.d, C.byref(need)), "ws")
ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
ref = None
for k in range(a.launches):
    check(lib.mfcd_recon_stats_tc(ptr(U), ptr(V), a.n, a.m, a.d, C.byref(xv), 1.0, ptr(ubar), ptr(vbar), ptr(stats), ptr(flag),
                                  ptr(ws), need.value, current_stream()), "tc")
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print(f"launch {k}: FAILED {str(e).splitlines()[0]}"); sys.exit(1)
    if int(flag.item()):
        print(f"launch {k}: timeout flag"); sys.exit(2)
    s = stats.clone()
    if ref is None: ref = s
    elif not torch.allclose(s[:, :6], ref[:, :6], rtol=1e-9, atol=0):
        print(f"launch {k}: results differ from launch 0 by {(s[:, :6] - ref[:, :6]).abs().max().item():.3e}")
print(f"ok: {a.launches} launches, d={a.d}")
