"""Per-step time of the reference-scale training loop (batch 64, deterministic mode) on synthetic triplets:
`python tools/small_step_probe.py [n m d steps]` prints us/step for mfcd_train_epoch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mfcd_b200
from mfcd_b200.store import TripletStore
from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch

n, m, d, steps = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1000, 1000, 10, 9375)))
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
N = steps * 64
store = TripletStore.from_columns(torch.from_numpy(rng.integers(0, n, N)), torch.from_numpy(rng.integers(0, m, N)),
                                  torch.from_numpy(rng.integers(0, m, N)),
                                  torch.from_numpy(rng.integers(0, 2, N).astype(np.float64)), device=dev)
torch.manual_seed(0)
model = MatrixFactorization(n, m, d)
fs = model.flat_state(dev)
spec = OptimizerSpec.adam(lr=1e-3, weight_decay=1e-5)
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    losses = run_epoch(fs, store, None, 64, spec, 1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"n={n} m={m} d={d}: {dt / steps * 1e6:.2f} us/step ({N / dt:.4g} triplets/s), mean loss {losses.mean().item():.6f}")
