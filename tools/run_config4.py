#!/usr/bin/env python
"""BASELINE.json config 4 as a JOB, through the product API: 100 000 users x 50 000 items, d = 64, ~1e9 unique
popularity-biased (zipf 1.5) triplets, data-parallel over the GPUs of one box (torchrun, one rank per GPU):

    ground truth (low-rank factors, never materialised: the dense X would be 20 GB)  ->  every rank samples, dedups
    and labels the triplets of ITS user range  ->  80/10/10 split  ->  `--epochs` epochs of training at a global batch
    of world x 2^22 (per-epoch device reshuffle + user grouping, K1 span kernel, fused K9 exchange)  ->  test loss /
    accuracy, ground-truth accuracy over all ranks' test shards.

Every phase is timed (wall clock, max over ranks after a barrier).  --triplets scales the job down for a smaller box
(default: 1.25e8 per GPU)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import structure
from mfcd_b200 import trainer as T

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000); ap.add_argument("--m", type=int, default=50_000)
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--triplets-per-gpu", type=float, default=1.25e8)
ap.add_argument("--batch-per-gpu", type=int, default=1 << 22)
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--lr", type=float, default=1e-3)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config4_job.json"))
a = ap.parse_args()

rank, world = T.dist_world(None)
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
times = {}


def phase(name, fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    times[name] = float(dt.item())
    if rank == 0:
        print(f"[config4] {name}: {times[name]:.3f} s", flush=True)
    return out


total = int(a.triplets_per_gpu) * world
p = 2.0 * total / (a.n * a.m)                       # run_experiment's num_triplets = n m p / 2
B = a.batch_per_gpu * world
base = structure._common_seed() if world > 1 else 12345
torch.manual_seed(base)
X = phase("generate_X (low-rank factors)", lambda: structure.generate_X(a.n, a.m, a.d, "cuda"))
if world > 1:
    torch.manual_seed(base + 1000003 * (rank + 1))
loaders = phase("sample + dedup + label + split (popularity)", lambda: structure.split_dataset_from_triplets(
    X, total, scale=1.0, K=1, batch_size=B, strategy="popularity", popularity_method="zipf", alpha=1.5,
    world_size=world))
train_loader, val_loader, test_loader = loaders
counts = torch.tensor([len(train_loader.store), len(val_loader.store), len(test_loader.store)], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(counts)
if world > 1:
    torch.manual_seed(base + 1)
model = structure.MatrixFactorization(a.n, a.m, a.d).to("cuda")
opt = torch.optim.Adam(model.parameters(), lr=a.lr, weight_decay=1e-5)
tl, vl = phase(f"train_model ({a.epochs} epoch(s), global batch {B})", lambda: T.train_model(
    model, train_loader, val_loader, opt, "cuda", num_epochs=a.epochs, mode="atomic", world_size=world))
test_loss, test_acc = phase("evaluate_model (test split)", lambda: structure.evaluate_model(model, test_loader, "cuda", world_size=world))
gt_loss, gt_acc = phase("compute_ground_truth_metrics", lambda: structure.compute_ground_truth_metrics(test_loader, X, "cuda", world_size=world))
if rank == 0:
    n_train = int(counts[0])
    res = {"config": vars(a), "world": world, "gpu": torch.cuda.get_device_name(0),
           "triplets": {"requested": total, "train": n_train, "val": int(counts[1]), "test": int(counts[2])},
           "global_batch": B, "steps_per_epoch": -(-n_train // B), "phases_s": times,
           "train_triplets_per_s": n_train * a.epochs / times[[k for k in times if k.startswith("train_model")][0]],
           "train_losses": tl, "val_losses": vl, "test_loss": test_loss, "test_accuracy": test_acc,
           "ground_truth_mse": gt_loss, "ground_truth_accuracy": gt_acc,
           "total_s": sum(times.values())}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k not in ("config",)})[:1500])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
