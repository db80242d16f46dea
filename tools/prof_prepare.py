#!/usr/bin/env python
"""Host-side profile of one repetition's preparation and finish at the Runs.ipynb cell-5 scale (1000 x 1000, d = 2,
p = 0.2, soft labels, 30 epochs): where does the caller thread of a concurrent sweep spend its time?"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import structure

def prep():
    return structure._prepare_rep(1000, 1000, 2, 0.2, 1.0, "cuda", 1e-3, 1e-5, 30, 1, "random", "zipf", 1.5, True, "base",
                                  64, None, 0, 1, record=True)

torch.manual_seed(0); np.random.seed(0)
P = prep(); structure._finish_rep(P, progress=False); torch.cuda.synchronize()      # warm-up
for name, fn in (("prepare", prep), ("finish", lambda: structure._finish_rep(prep(), progress=False))):
    torch.cuda.synchronize(); t = time.perf_counter()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    pr.disable()
    print(f"== {name}: {(time.perf_counter() - t) / 5 * 1e3:.1f} ms per call")
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
