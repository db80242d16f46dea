#!/usr/bin/env python
"""BASELINE.json config 5: the sampling-strategy x d sweep at 20,000 x 20,000 (random / margin / svd / popularity,
d in {2, 8, 32, 128}) with the full evaluation (accuracy, reconstruction error, Pearson / Spearman, svd error),
through structure.run_experiment, with a per-function wall-clock breakdown.

p = 0.1 -> 2e7 unique triplets per experiment (16 M training samples).  The reference's batch of 64 would mean
250 000 dense-Adam steps per epoch; the sweep uses the batch_size knob (default 65536) and `--epochs` epochs."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import structure

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=20000); ap.add_argument("--m", type=int, default=20000)
ap.add_argument("--p", type=float, default=0.1)
ap.add_argument("--ds", default="2,8,32,128"); ap.add_argument("--strategies", default="random,margin,svd,popularity")
ap.add_argument("--epochs", type=int, default=10); ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--lr", type=float, default=1e-2)
ap.add_argument("--wd", type=float, default=None,
                help="weight decay; default 1e-5 * 64 / batch: the reference's coupled L2 (1e-5 at batch 64) acts on a "
                     "batch-MEAN gradient, so at batch 65536 an unscaled 1e-5 outweighs the per-row data gradient "
                     "(~1e-5 |V|) and every uniform-item run collapses to U = V = 0 (loss ln 2) -- plain torch on the "
                     "CPU does the same, tools/config5_dynamics_cpu.py")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5.json"))
a = ap.parse_args()
if a.wd is None:
    a.wd = 1e-5 * 64.0 / a.batch

FUNCS = ("generate_X", "split_dataset_from_triplets", "evaluate_model", "compute_reconstruction_error",
         "compute_alpha_and_norm_ratios", "compute_ground_truth_metrics")
store = {}


def timed(mod, name):
    orig = getattr(mod, name)

    def wrap(*args, **kw):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = orig(*args, **kw)
        torch.cuda.synchronize(); store[name] = store.get(name, 0.0) + time.perf_counter() - t
        return out
    setattr(mod, name, wrap)


for f in FUNCS:
    timed(structure, f)
timed(structure._trainer, "train_model")
rows = []
torch.manual_seed(0); np.random.seed(0)
structure.run_experiment(200, 200, 4, 0.2, 1.0, "cuda", 1e-2, 1e-5, reps=1, num_epochs=1, batch_size=1024)   # warm-up
for strategy in a.strategies.split(","):
    for d in [int(x) for x in a.ds.split(",")]:
        store.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        try:
            out = structure.run_experiment(a.n, a.m, d, a.p, 1.0, "cuda", a.lr, a.wd, reps=1, num_epochs=a.epochs,
                                           K=1, strategy=strategy, batch_size=a.batch)
            torch.cuda.synchronize(); total = time.perf_counter() - t0
            n_train = int(0.8 * int(a.n * a.m * a.p / 2))
            row = {"strategy": strategy, "d": d, "total_s": total, "breakdown_s": dict(store),
                   "train_triplets_per_s": n_train * a.epochs / store["train_model"],
                   "accuracy": out["accuracy"][0], "gt_accuracy": out["gt_accuracy"][0],
                   "reconstruction_error": out["reconstruction_errors"][0], "pearson": out["pearson_corr"][0],
                   "spearman": out["spearman_corr"][0], "svd_error_scaled": out["svd_error_scaled"][0],
                   "alpha": out["alpha"][0], "train_loss_first_last": [out["train_losses"][0][0], out["train_losses"][0][-1]],
                   "val_loss_last": out["val_losses"][0][-1]}
        except Exception as e:        # e.g. the reference's own svds() rank error for some shapes: record, go on
            row = {"strategy": strategy, "d": d, "error": repr(e)[:300]}
        rows.append(row)
        print(json.dumps(row), flush=True)
        torch.cuda.empty_cache()
res = {"config": vars(a), "gpu": torch.cuda.get_device_name(0), "rows": rows}
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
