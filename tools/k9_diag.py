"""K9 alone (no K1, so no skew between the ranks' compute): device time per exchange, for the in-kernel-flag
variants (double-buffered / owner-zeroes, p2p / multimem), NCCL all-reduce + K3, and K3 alone.  Run under torchrun."""
import os, sys, time, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mfcd_b200
from mfcd_b200._lib import lib, check, ptr, current_stream
from mfcd_b200 import dist as mdist
from mfcd_b200.trainer import OptimizerSpec, _FlatState

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, m, d = 100_000, 50_000, 64
spec = OptimizerSpec.adam(lr=1e-3, weight_decay=1e-5)

def timed(fn, iters=40, warm=8):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

res = {}
for name, env in (("k9 double-buffered, auto", {}), ("k9 double-buffered, p2p", {"MFCD_DP_MULTIMEM": "off"}),
                  ("k9 double-buffered, multimem", {"MFCD_DP_MULTIMEM": "on"}),
                  ("k9 owner-zeroes, auto", {"MFCD_DP_DOUBLE_BUFFER": "0"})):
    os.environ.pop("MFCD_DP_MULTIMEM", None); os.environ.pop("MFCD_DP_DOUBLE_BUFFER", None)
    os.environ.update(env)
    ex = mdist.PeerExchange((n + m) * d, dev)
    fs = _FlatState(n, m, d, dev, params=ex.params, grads=ex.grads)
    step = [0]
    def k9():
        step[0] += 1
        ex.step(fs, spec, step[0])
    res[name + (" [multimem]" if ex.multimem else " [p2p]")] = timed(k9)
    ex.check_error()
    del ex, fs
g = torch.randn((n + m) * d, device=dev); p = torch.randn_like(g); m1 = torch.zeros_like(g); v1 = torch.zeros_like(g)
def adam():
    check(lib.mfcd_adam_update(ptr(p), ptr(g), ptr(m1), ptr(v1), g.numel(), 1e-3, 0.9, 0.999, 1e-8, 1e-5, 1, 1, current_stream()), "adam")
def nccl():
    dist.all_reduce(g); adam()
res["nccl all-reduce + K3"] = timed(nccl)
res["K3 alone"] = timed(adam)
if rank == 0:
    print(json.dumps({"world": world, "ms_per_exchange": res}, indent=1))
dist.destroy_process_group()
