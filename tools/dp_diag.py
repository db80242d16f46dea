"""Timing breakdown of the data-parallel step (run under torchrun)."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfcd_b200
from mfcd_b200._lib import lib, check, ptr, current_stream

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
numel = 150_000 * 64
g = torch.randn(numel, device=dev); p = torch.randn(numel, device=dev); m = torch.zeros_like(p); v = torch.zeros_like(p)

def timed(fn, iters=50, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    return a.elapsed_time(b) / iters, (t1 - t0) * 1e3 / iters

def adam(a=0, b=numel, step=1):
    check(lib.mfcd_adam_update(ptr(p[a:]), ptr(g[a:]), ptr(m[a:]), ptr(v[a:]), b - a, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step, 0, current_stream()), "adam")

res = {}
res["allreduce_whole"] = timed(lambda: dist.all_reduce(g))
for mb in (4, 16):
    be = mb * (1 << 20) // 4
    bounds = [(a, min(a + be, numel)) for a in range(0, numel, be)]
    def bucketed():
        works = [dist.all_reduce(g[a:b], async_op=True) for a, b in bounds]
        for (a, b), w in zip(bounds, works):
            w.wait(); adam(a, b)
    res[f"bucketed_{mb}MB+adam"] = timed(bucketed)
def whole_then_adam():
    dist.all_reduce(g); adam()
res["whole+adam"] = timed(whole_then_adam)
res["adam_only"] = timed(adam)
if rank == 0:
    for k, (dev_ms, wall_ms) in res.items():
        print(f"{k:24s} device {dev_ms:8.4f} ms   wall {wall_ms:8.4f} ms")
dist.destroy_process_group()
