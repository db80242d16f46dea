#!/usr/bin/env python
"""Sweep-level concurrency: a Runs.ipynb cell-5-like sweep (1000 x 1000, d = 2, p = 0.2,
soft labels, 30 epochs; a subset of its weight_decay x K x s grid, `reps` repetitions each) run sequentially and with
k repetitions in flight on one GPU.  Prints wall times, the speed-up and whether the results are identical."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import structure

ap = argparse.ArgumentParser()
ap.add_argument("--concurrency", default="4,8,16")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--epochs", type=int, default=30)
ap.add_argument("--devices", default=None)
a = ap.parse_args()
grid = dict(n=1000, m=1000, d=2, p=0.2, lr=1e-3, weight_decay=[1e-6, 1e-5, 1e-4], num_epochs=a.epochs, reps=a.reps,
            s=[0.1, 1.0], K=[1, 4], device="cuda", soft_label=True)


def same(x, y):
    if isinstance(x, dict):
        return x.keys() == y.keys() and all(same(x[k], y[k]) for k in x)
    if isinstance(x, (list, tuple)):
        return len(x) == len(y) and all(same(p, q) for p, q in zip(x, y))
    if isinstance(x, np.ndarray):
        return np.array_equal(x, y)
    return x == y or (x != x and y != y)


def run(conc, devices=None):
    torch.manual_seed(0); np.random.seed(0)
    torch.cuda.synchronize(); t = time.perf_counter()
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        res = structure.parameter_scan(**grid, concurrency=conc, devices=devices)
    torch.cuda.synchronize()
    return time.perf_counter() - t, res

run(None)                                         # warm-up (lazy CUDA / cuSOLVER init)
t_seq, r_seq = run(None)
units = len(r_seq) * a.reps
out = {"grid": {k: v for k, v in grid.items()}, "experiments": len(r_seq), "repetitions": units,
       "sequential_s": t_seq, "per_repetition_s": t_seq / units, "concurrent": {}}
def trace_summary(path):
    """where the wall time of a concurrent sweep goes: the caller's preparation, and the workers' phases"""
    tr = json.load(open(path))
    agg = {}
    for what, _unit, _thr, t0, t1 in tr["units"]:
        agg.setdefault(what, []).append(t1 - t0)
    for what, _thr, t0, t1 in tr["phases"]:
        agg.setdefault("phase:" + what, []).append(t1 - t0)
    return {k: {"count": len(v), "total_s": sum(v), "mean_s": sum(v) / len(v)} for k, v in agg.items()}


for c in [int(x) for x in a.concurrency.split(",")]:
    trace_path = os.path.join(ROOT, "gpurun_out", f"sweep_trace_c{c}.json")
    os.makedirs(os.path.dirname(trace_path), exist_ok=True)
    os.environ["MFCD_SWEEP_TRACE"] = trace_path
    t, r = run(c)
    del os.environ["MFCD_SWEEP_TRACE"]
    out["concurrent"][str(c)] = {"wall_s": t, "speedup": t_seq / t, "identical_to_sequential": same(r, r_seq),
                                 "where": trace_summary(trace_path)}
if a.devices:
    t, r = run(8 * torch.cuda.device_count(), devices="all")
    out["all_gpus"] = {"gpus": torch.cuda.device_count(), "wall_s": t, "speedup": t_seq / t, "identical_to_sequential": same(r, r_seq)}
print(json.dumps(out, indent=1))
