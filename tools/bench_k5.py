"""Times the reconstruction-statistics pass (K5) with both engines and prints X-stream GB/s."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mfcd_b200
from mfcd_b200 import metrics
from mfcd_b200.store import GroundTruth
from mfcd_b200.trainer import MatrixFactorization

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192); ap.add_argument("--m", type=int, default=20480)
ap.add_argument("--d", type=int, default=64); ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--engines", default="simt,tc")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MatrixFactorization(a.n, a.m, a.d)
X = torch.randn(a.n, a.m, device=dev)
gt = GroundTruth(X=X)
peak = 6535.7
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"n": a.n, "m": a.m, "d": a.d, "x_bytes": 4 * a.n * a.m, "flops": 2 * a.n * a.m * a.d}
ref = None
for eng in a.engines.split(","):
    metrics._row_stats(model, gt, 1.0, engine=eng)          # warm-up (includes host copy of the stats)
    fs = model.flat_state(dev)
    import ctypes as C
    from mfcd_b200._lib import lib, check, ptr, current_stream
    ubar = torch.zeros(a.d, device=dev); vbar = torch.zeros(a.d, device=dev)
    stats = torch.empty((a.n, 8), dtype=torch.float64, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    xv = gt.xview()
    need = C.c_size_t(0)
    check(lib.mfcd_recon_stats_tc_workspace_bytes(a.n, a.m, a.d, C.byref(need)), "ws")
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
    def run():
        if eng == "tc":
            check(lib.mfcd_recon_stats_tc(ptr(fs.U), ptr(fs.V), a.n, a.m, a.d, C.byref(xv), 1.0, ptr(ubar), ptr(vbar), ptr(stats), ptr(flag), ptr(ws), need.value, current_stream()), "tc")
        else:
            check(lib.mfcd_recon_stats(ptr(fs.U), ptr(fs.V), a.n, a.m, a.d, C.byref(xv), 1.0, ptr(ubar), ptr(vbar), ptr(stats), current_stream()), "simt")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    s = stats.cpu().numpy()
    if ref is None: ref = s
    out[eng] = {"ms": ms, "x_stream_GBps": out["x_bytes"] / ms / 1e6, "frac_of_hbm_peak": out["x_bytes"] / ms / 1e6 / peak,
                "tflops": out["flops"] / ms / 1e9, "max_rel_diff_vs_first": float(np.abs(s[:, [1, 3, 4, 5]] - ref[:, [1, 3, 4, 5]]).max() / np.abs(ref[:, [1, 3, 4, 5]]).max()),
                "timeout_flag": int(flag.item())}
print(json.dumps(out))
