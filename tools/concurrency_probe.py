#!/usr/bin/env python
"""Why does a short kernel wait for another stream's epoch kernel?  Times small operations on the caller's stream while a
worker stream runs (a) the persistent epoch kernel, (b) torch.cuda._sleep; caller on the legacy default stream or on
its own stream."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import structure
from mfcd_b200 import trainer as T

dev = torch.device("cuda", 0)
print("CUDA_DEVICE_MAX_CONNECTIONS =", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
torch.manual_seed(0); np.random.seed(0)
P = structure._prepare_rep(1000, 1000, 2, 0.2, 1.0, "cuda", 1e-3, 1e-5, 30, 1, "random", "zipf", 1.5, True, "base",
                           64, None, 0, 1, record=False)
model, (train_loader, val_loader, test_loader) = P["model"], P["loaders"]
fs = model.flat_state(dev)
spec = T.OptimizerSpec(P["optimizer"])
store = train_loader.store
perm = torch.randperm(len(store), device=dev, dtype=torch.int32)
T.run_epoch(fs, store, perm, 64, spec, T.MODE_DETERMINISTIC); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); T.run_epoch(fs, store, perm, 64, spec, T.MODE_DETERMINISTIC); e1.record(); torch.cuda.synchronize()
print("one epoch kernel: %.2f ms" % e0.elapsed_time(e1))

def small_ops(tag, stream):
    lat = []
    with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.default_stream()):
        for _ in range(20):
            t = time.perf_counter()
            x = torch.randperm(80000, device=dev)
            (stream or torch.cuda.default_stream()).synchronize()
            lat.append((time.perf_counter() - t) * 1e3)
    print(f"  {tag}: small-op latency ms: median {np.median(lat):.3f} max {max(lat):.3f}")

def busy_epochs(stream, n):
    with torch.cuda.stream(stream):
        for _ in range(n):
            T.run_epoch(fs, store, perm, 64, spec, T.MODE_DETERMINISTIC)

def busy_sleep(stream, n):
    with torch.cuda.stream(stream):
        for _ in range(n):
            torch.cuda._sleep(int(7e-3 * 1.9e9))

own = torch.cuda.Stream()
small_ops("idle GPU, default stream", None)
small_ops("idle GPU, own stream", own)
for name, busy in (("epoch kernels", busy_epochs), ("_sleep kernels", busy_sleep)):
    for tag, st in (("default stream", None), ("own stream", own)):
        w = torch.cuda.Stream()
        th = threading.Thread(target=busy, args=(w, 40)); th.start(); time.sleep(0.02)
        small_ops(f"{name} on a worker stream; caller on {tag}", st)
        th.join(); torch.cuda.synchronize()
