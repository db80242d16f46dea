"""Wall-clock breakdown of structure.run_experiment at the reference's own scales (configs 1 and 2)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import structure

def timed(fn_name, store):
    # run_experiment calls the trainer's train_model directly; everything else through the module's own names
    mod = structure._trainer if fn_name == "train_model" else structure
    orig = getattr(mod, fn_name)
    def wrap(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = orig(*a, **k)
        torch.cuda.synchronize(); store[fn_name] = store.get(fn_name, 0.0) + time.perf_counter() - t
        return out
    setattr(mod, fn_name, wrap)
    return orig

res = {}
for tag, (n, m, d, p, s, K, epochs) in {"c1": (100, 100, 2, 0.1, 1.0, 1, 30), "c2": (1000, 1000, 10, 0.5, 1.0, 3, 5),
                                         "runs_ipynb_cell3": (1000, 1000, 2, 0.2, 1.0, 1, 30)}.items():
    store = {}
    origs = {f: timed(f, store) for f in ("generate_X", "split_dataset_from_triplets", "train_model", "evaluate_model",
                                         "compute_reconstruction_error", "compute_alpha_and_norm_ratios",
                                         "compute_ground_truth_metrics")}
    torch.manual_seed(0); np.random.seed(0)
    structure.run_experiment(n, m, d, p, s, "cuda", 1e-3, 1e-5, reps=1, num_epochs=1, K=K)      # warm-up (lazy init)
    store.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = structure.run_experiment(n, m, d, p, s, "cuda", 1e-3, 1e-5, reps=1, num_epochs=epochs, K=K)
    torch.cuda.synchronize(); total = time.perf_counter() - t0
    for f, o in origs.items(): setattr(structure._trainer if f == "train_model" else structure, f, o)
    n_train = int(0.8 * int(n * m * p / 2)) * K
    res[tag] = {"config": dict(n=n, m=m, d=d, p=p, K=K, epochs=epochs), "total_s": total, "breakdown_s": store,
                "train_samples_per_epoch": n_train, "steps_per_epoch": (n_train + 63) // 64,
                "train_triplets_per_s": n_train * epochs / store["train_model"],
                "final_train_loss": out["train_losses"][0][-1], "accuracy": out["accuracy"][0]}
print(json.dumps(res, indent=1))
