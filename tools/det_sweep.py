#!/usr/bin/env python
"""Deterministic large-batch mode: the two engines (fixed-point integer atomics / sort + segmented reduction) timed
over batch sizes at the config-3 and config-4 table shapes, uniform and zipf(1.5) items.  us per K1 call."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mfcd_b200._lib import lib, check, ptr, current_stream

dev = torch.device("cuda", 0)
out = []
for tag, n, m, d in (("c3", 10_000, 5_000, 32), ("c4", 100_000, 50_000, 64)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    U = torch.randn(n, d, device=dev, generator=g) / d ** 0.5; V = torch.randn(m, d, device=dev, generator=g) / d ** 0.5
    gU = torch.zeros_like(U); gV = torch.zeros_like(V); loss = torch.zeros(1, device=dev)
    pr = 1.0 / torch.arange(1, m + 1, device=dev, dtype=torch.float64) ** 1.5
    for dist in ("uniform", "zipf"):
        for B in (1024, 4096, 16384, 65536, 262144, 1048576):
            rec = torch.empty((B, 4), dtype=torch.int32, device=dev)
            rec[:, 0] = torch.randint(0, n, (B,), generator=g, device=dev)
            if dist == "uniform":
                rec[:, 1] = torch.randint(0, m, (B,), generator=g, device=dev); rec[:, 2] = torch.randint(0, m, (B,), generator=g, device=dev)
            else:
                rec[:, 1] = torch.multinomial(pr, B, replacement=True, generator=g).int(); rec[:, 2] = torch.multinomial(pr, B, replacement=True, generator=g).int()
            rec[:, 3] = torch.randint(0, 2, (B,), generator=g, device=dev).float().view(torch.int32)
            row = {"shape": tag, "items": dist, "B": B}
            for eng in ("fixed", "sort"):
                need = C.c_size_t(0)
                if eng == "fixed":
                    check(lib.mfcd_det_fixed_workspace_bytes(d, n, m, C.byref(need)), "ws"); fn = lib.mfcd_triplet_fwd_bwd_det_fixed
                else:
                    check(lib.mfcd_det_workspace_bytes(B, d, C.byref(need)), "ws"); fn = lib.mfcd_triplet_fwd_bwd_det_sort
                ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
                def run():
                    check(fn(ptr(U), ptr(V), ptr(rec), None, 0, B, d, 1.0 / B, n, m, ptr(gU), ptr(gV), ptr(loss), ptr(ws),
                             need.value, current_stream()), eng)
                for _ in range(3): run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 20 if B <= 65536 else 5
                e0.record()
                for _ in range(iters): run()
                e1.record(); torch.cuda.synchronize()
                row[eng + "_us"] = round(e0.elapsed_time(e1) / iters * 1e3, 1)
            out.append(row); print(json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "det_sweep.json"), "w"), indent=1)
