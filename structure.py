"""Drop-in for the reference's ``structure.py``: same public names, signatures,
return types and error behaviour, so ``Runs.ipynb`` drives it unchanged
(``from structure import parameter_scan`` etc.).  The hot path -- labels, split,
training loop, BCE loss, evaluation, reconstruction metrics -- runs in the
hand-written sm_100a kernels of ``mfcd_b200`` (C ABI: include/mfcd_b200.h).

Where the reference does the work: /root/reference/structure.py
  parameter_scan :81-255, run_experiment :306-450, BTLPreferenceDataset :465-531,
  get_triplets_from_X :533-588, generate_X :590-663, split_dataset_from_triplets
  :666-742, MatrixFactorization :746-795, train_model :812-878, evaluate_model
  :881-921, compute_reconstruction_error :925-955, compute_alpha_and_norm_ratios
  :958-1082, compute_ground_truth_metrics :1085-1127, evaluate_ground_truth
  :1154-1200, parameter_scan_ground_truth :1203-1269.

``device`` keeps its meaning for user-visible tensors; the compute device is
always a CUDA GPU (there is no CPU path: without one every call raises).
Extra knobs are keyword-only with reference defaults.
"""
import os

os.environ.setdefault("OMP_NUM_THREADS", "4")      # the reference pins this at import (structure.py:3)

import itertools
import time
import types
import pickle

import numpy as np
import torch
from tqdm import tqdm

import mfcd_b200  # noqa: F401
from mfcd_b200 import config as _cfg
from mfcd_b200 import metrics as _metrics
from mfcd_b200 import sampling as _sampling
from mfcd_b200 import trainer as _trainer
from mfcd_b200 import dist as _mdist
from mfcd_b200.store import GroundTruth, TripletLoader, TripletStore, compute_device
from mfcd_b200.trainer import MatrixFactorization  # noqa: F401  (structure.py:746)
from generation_data import *  # noqa: F401,F403  (the reference re-exports it, structure.py:17)
import generation_data as _gen

_SCAN_KEYS = ("n", "m", "d", "p", "lr", "weight_decay", "num_epochs", "reps", "s", "K", "d1", "strategy",
              "popularity_method", "alpha", "soft_label", "generation")


# ---------------------------------------------------------------------------
# sweeps
# ---------------------------------------------------------------------------
def _plain(x):
    """numpy scalars -> python scalars (structure.py:128-134)."""
    if isinstance(x, (np.float32, np.float64)):
        return float(x)
    if isinstance(x, np.integer):
        return int(x)
    return x


def _normalise_grid(params):
    """-> (dict name -> list of values, the list-valued entries, lists_are_synchronised)"""
    norm = {}
    for name, value in params.items():
        if isinstance(value, np.ndarray):
            norm[name] = list(value)
        elif isinstance(value, list):
            norm[name] = [_plain(x) for x in value]
        else:
            norm[name] = _plain(value)
    lists = [v for v in norm.values() if isinstance(v, list)]
    synchronised = len(lists) <= 1 or all(len(v) == len(lists[0]) for v in lists)
    for name, value in norm.items():
        if not isinstance(value, (list, tuple)):
            norm[name] = [value]
    return norm, lists, synchronised


def _append_pickle(path, new_items):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    stored = []
    if os.path.exists(path):
        with open(path, "rb") as f:
            stored = pickle.load(f)
    stored.extend(new_items)
    with open(path, "wb") as f:
        pickle.dump(stored, f)
    print(f"✅ Saved {len(new_items)} new experiments to {path}")


def parameter_scan(n=1000, m=1000, d=2, p=0.5, s=1.0, device='cpu',
                   lr=1e-3, weight_decay=1e-5, num_epochs=30, reps=1, strategy="random",
                   open_browser=False, linear=False, K=1, d1=None,
                   save_path=None, save_every=None, popularity_method="zipf",
                   alpha=1.5, soft_label=False, generation="base", *,
                   batch_size=64, mode=None, world_size=None, concurrency=None, devices=None):
    """Grid (default) or synchronised linear sweep over scalar-or-list arguments;
    one ``run_experiment`` per configuration.  Returns ``[{'params', 'results'}]``
    -- or ``[]`` when ``save_path`` is set, because saved chunks are dropped from
    memory exactly like the reference does (structure.py:186, :200-202).  An
    existing ``save_path`` is deleted first (:151-153).

    Keyword-only extras (reference behaviour by default): ``batch_size`` / ``mode`` / ``world_size`` are forwarded to
    run_experiment; ``concurrency`` = k > 1 keeps up to k repetitions in flight on separate CUDA streams (default: env
    MFCD_SWEEP_CONCURRENCY, else sequential), ``devices`` = "all" or a list of CUDA devices spreads them over the
    GPUs of the box.  Results are identical to the sequential sweep and come back in the same order."""
    if concurrency is None and os.environ.get("MFCD_SWEEP_CONCURRENCY"):
        concurrency = int(os.environ["MFCD_SWEEP_CONCURRENCY"])
    grid, lists, synchronised = _normalise_grid(dict(zip(_SCAN_KEYS, (
        n, m, d, p, lr, weight_decay, num_epochs, reps, s, K, d1, strategy, popularity_method, alpha,
        soft_label, generation))))

    if save_path and os.path.exists(save_path):
        print(f"🧹 Removing existing file at {save_path}")
        os.remove(save_path)

    if not linear:
        configs = [dict(zip(grid.keys(), combo)) for combo in itertools.product(*grid.values())]
    elif synchronised:
        configs = [{k: (v[idx] if len(v) > 1 else v[0]) for k, v in grid.items()}
                   for idx in range(len(lists[0]))]
    else:
        raise ValueError("The linear scan is not possible because the parameters are not synchronized.")

    if concurrency is not None and int(concurrency) > 1 and _trainer.dist_world(world_size)[1] == 1:
        return _scan_concurrent(configs, device, open_browser, batch_size, mode, int(concurrency), devices,
                                save_path, save_every)
    pending = []
    for cfg in configs:
        print(f"\nRunning experiment with parameters: {cfg}")
        results = run_experiment(
            n=cfg['n'], m=cfg['m'], d=cfg['d'], p=cfg['p'], s=cfg['s'], device=device, lr=cfg['lr'],
            weight_decay=cfg['weight_decay'], reps=cfg['reps'], num_epochs=cfg['num_epochs'],
            open_browser=open_browser, K=cfg['K'], d1=cfg['d1'], strategy=cfg['strategy'],
            popularity_method=cfg['popularity_method'], alpha=cfg['alpha'], soft_label=cfg['soft_label'],
            generation=cfg['generation'], batch_size=batch_size, mode=mode, world_size=world_size)
        pending.append({'params': cfg, 'results': results})
        if save_path and save_every and len(pending) >= save_every:
            _append_pickle(save_path, pending)
            pending = []
    if save_path and pending:
        _append_pickle(save_path, pending)
        pending = []
    return pending


def _adopt_on_stream(obj, stream, seen=None, depth=0):
    """Tell the caching allocator that every CUDA tensor reachable from ``obj`` is used on ``stream``.

    A repetition is prepared on the caller's stream and finished on a worker stream.  Tensors made during the
    preparation (ground truth, records, the recorded per-epoch permutations, the tables) belong to the CALLER's
    stream pool: when the worker drops one of them mid-run (a consumed epoch permutation, a re-ordered store), the
    allocator would hand the block to the next preparation while the worker's kernels are still reading it.
    ``Tensor.record_stream`` defers that reuse until the worker stream has passed the point of release."""
    if seen is None:
        seen = set()
    if id(obj) in seen or depth > 12:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
        return
    if obj is None or isinstance(obj, (str, bytes, int, float, bool, complex, np.ndarray, np.generic, torch.device,
                                       torch.dtype, torch.cuda.Stream, torch.cuda.Event, type, types.ModuleType,
                                       types.FunctionType, types.MethodType, types.BuiltinFunctionType)):
        return
    if isinstance(obj, dict):
        children = list(obj.keys()) + list(obj.values())
    elif isinstance(obj, (list, tuple, set, frozenset)) or type(obj).__name__ == "deque":
        children = list(obj)
    elif hasattr(obj, "__dict__"):
        children = list(vars(obj).values())
    else:
        return
    for child in children:
        _adopt_on_stream(child, stream, seen, depth + 1)


def _scan_concurrent(configs, device, open_browser, batch_size, mode, concurrency, devices, save_path, save_every):
    """The sweep with up to ``concurrency`` repetitions in flight.

    A reference-sized experiment leaves most of a B200 idle (its persistent epoch kernel runs on ~10 of 148 SMs), and
    the reference's sweeps are thousands of such experiments.  Every repetition is PREPARED on the calling thread, in
    the sequential order -- that is where the host generators are consumed (ground truth, triplets, labels, split,
    model initialisation, and, recorded ahead of time, the per-epoch loader draws) -- and its GPU work (training,
    evaluation, metrics) runs on a worker thread with its own CUDA stream, round-robin over ``devices``.  Results
    come back in the reference's order and are IDENTICAL to a sequential run under the same seeds."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    if devices is None:
        devs = [compute_device(device)]
    elif devices == "all":
        devs = [torch.device("cuda", k) for k in range(torch.cuda.device_count())]
    else:
        devs = [torch.device(x) for x in devices]
    slots = threading.Semaphore(concurrency)
    local = threading.local()
    trace = [] if os.environ.get("MFCD_SWEEP_TRACE") else None        # (what, unit, thread, t0, t1) wall-clock events
    global _PHASE_TRACE
    _PHASE_TRACE = [] if trace is not None else None

    def work(P, dev, ready, is_last, unit_id=0):
        t_begin = time.perf_counter()
        try:
            with torch.cuda.device(dev):
                streams = local.__dict__.setdefault("streams", {})
                if dev not in streams:
                    streams[dev] = torch.cuda.Stream(device=dev)
                st = streams[dev]
                st.wait_event(ready)                       # the preparation ran on the caller's stream
                _adopt_on_stream(P, st)
                with torch.cuda.stream(st):
                    values = _finish_rep(P, is_last=is_last, open_browser=open_browser, progress=False)
                    st.synchronize()
                return values
        finally:
            if trace is not None:
                trace.append(("finish", unit_id, threading.get_ident(), t_begin, time.perf_counter()))
            slots.release()

    jobs = []                                              # (cfg, [future per repetition])
    unit = 0
    with ThreadPoolExecutor(max_workers=concurrency) as pool:
        pending, done_upto = [], 0

        def flush(block):
            """move finished experiments, in order, to `pending`; save chunks like the sequential loop does"""
            nonlocal done_upto, pending
            while done_upto < len(jobs):
                cfg, futs = jobs[done_upto]
                if not block and not all(f.done() for f in futs):
                    break
                out = {key: [] for key in _RESULT_KEYS}
                for f in futs:
                    values = f.result()
                    for key in _RESULT_KEYS:
                        out[key].append(values[key])
                pending.append({'params': cfg, 'results': out})
                jobs[done_upto] = (cfg, [])
                done_upto += 1
                if save_path and save_every and len(pending) >= save_every:
                    _append_pickle(save_path, pending)
                    pending = []

        for cfg in configs:
            print(f"\nRunning experiment with parameters: {cfg}")
            futs = []
            for rep in range(cfg['reps']):
                t_wait = time.perf_counter()
                slots.acquire()
                dev = devs[unit % len(devs)]
                unit += 1
                t_prep = time.perf_counter()
                try:
                    with torch.cuda.device(dev):
                        dev_arg = device if len(devs) == 1 else dev
                        P = _prepare_rep(cfg['n'], cfg['m'], cfg['d'], cfg['p'], cfg['s'], dev_arg, cfg['lr'],
                                         cfg['weight_decay'], cfg['num_epochs'], cfg['K'], cfg['strategy'],
                                         cfg['popularity_method'], cfg['alpha'], cfg['soft_label'], cfg['generation'],
                                         batch_size, mode, 0, 1, record=True)
                        ready = torch.cuda.Event()
                        ready.record(torch.cuda.current_stream(dev))
                except BaseException:
                    slots.release()
                    raise
                if trace is not None:
                    trace.append(("wait_slot", unit, 0, t_wait, t_prep))
                    trace.append(("prepare", unit, 0, t_prep, time.perf_counter()))
                futs.append(pool.submit(work, P, dev, ready, rep == cfg['reps'] - 1, unit))
            jobs.append((cfg, futs))
            flush(block=False)
        flush(block=True)
    if trace is not None:
        import json
        with open(os.environ["MFCD_SWEEP_TRACE"], "w") as f:
            json.dump({"units": sorted(trace, key=lambda e: e[3]), "phases": sorted(_PHASE_TRACE, key=lambda e: e[2])}, f)
        _PHASE_TRACE = None
    if save_path and pending:
        _append_pickle(save_path, pending)
        pending = []
    return pending


def print_return_structure_types(obj, prefix="root"):
    """Debug helper: print the type skeleton of a nested result (structure.py:258-302)."""
    if isinstance(obj, dict):
        for key, value in obj.items():
            print_return_structure_types(value, f"{prefix}.{key}")
    elif isinstance(obj, (list, tuple)):
        kinds = {type(el).__name__ for el in obj}
        inner = "empty" if not obj else (kinds.pop() if len(kinds) == 1 else "mixed")
        print(f"{prefix}: {type(obj).__name__}[{inner}]")
    elif isinstance(obj, torch.Tensor):
        print(f"{prefix}: torch.Tensor")
    else:
        print(f"{prefix}: {type(obj).__name__}")


def _common_seed():
    """one 62-bit seed, the same on every rank of the data-parallel job (rank 0's draw)"""
    import torch.distributed as dist
    seed = torch.tensor([_sampling.fresh_seed()], dtype=torch.int64, device=compute_device(None))
    dist.broadcast(seed, src=0)
    return int(seed.item()) % (2 ** 62)


_RESULT_KEYS = (
    "reconstruction_errors", "log_likelihoods", "accuracy", "gt_log_likelihoods", "gt_accuracy",
    "train_losses", "val_losses", "alpha", "norm_X", "norm_ratio", "reconstruction_error_scaled",
    "pearson_corr", "pearson_std", "spearman_corr", "spearman_std", "svd_error_scaled", "slopes",
    "pearson_corr_matrix", "spearman_corr_matrix", "reconstruction_error_scaled_per_row", "alpha_per_row",
    "sampled_UVT_rows", "sampled_X_rows")


def _prepare_rep(n, m, d, p, s, device, lr, weight_decay, num_epochs, K, strategy, popularity_method, alpha,
                 soft_label, generation, batch_size, mode, rank, world, record):
    """The part of one repetition that consumes the host generators: ground truth, triplets, labels, split, model
    initialisation (structure.py:353-364).  record=True additionally makes, now and in the sequential order, the
    draws the rest of the repetition would make (per-epoch loader seeds / permutations, the two sampled row
    indices), so the GPU work can run later -- concurrently with other experiments -- on the same random streams."""
    if world > 1:
        base = _common_seed()
        torch.manual_seed(base)                      # the same ground truth on every rank
    X = generate_X(n, m, d, device, generation=generation)
    num_triplets = int(n * m * p / 2)
    if world > 1:
        torch.manual_seed(base + 1000003 * (rank + 1))   # per-rank sampling / label streams
    train_loader, val_loader, test_loader = split_dataset_from_triplets(
        X, num_triplets, scale=s, K=K, batch_size=batch_size, strategy=strategy,
        popularity_method=popularity_method, alpha=alpha, soft_label=soft_label, world_size=world)
    if world > 1:
        torch.manual_seed(base + 1)                  # common stream again: model init, sampled rows
    model = MatrixFactorization(n, m, d).to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)
    P = dict(X=X, s=s, device=device, loaders=(train_loader, val_loader, test_loader), model=model,
             optimizer=optimizer, num_epochs=num_epochs, mode=mode, world=world, rand_indices=None)
    if record:
        atomic = _trainer.resolve_mode(_cfg.SCATTER_MODE if mode is None else mode, batch_size) == _trainer.MODE_ATOMIC
        for _ in range(num_epochs):                  # train_model: one training and one validation iterator per epoch
            train_loader.record_draws(1, batch_size, atomic)
            val_loader.record_draws(1)
        test_loader.record_draws(1)                  # evaluate_model
        P["rand_indices"] = torch.randperm(X.shape[0])[:2]
        test_loader.record_draws(1)                  # compute_ground_truth_metrics
    return P


_PHASE_TRACE = None        # list: (phase, thread id, t0, t1) of every _finish_rep phase, when a sweep trace is on


class _phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        self.t0 = time.perf_counter() if _PHASE_TRACE is not None else 0.0

    def __exit__(self, *exc):
        if _PHASE_TRACE is not None:
            import threading
            _PHASE_TRACE.append((self.name, threading.get_ident(), self.t0, time.perf_counter()))
        return False


def _finish_rep(P, is_last=False, open_browser=False, progress=True):
    """Train, evaluate and measure one prepared repetition (structure.py:368-409) -> {result key: value}."""
    X, s, device, model, world = P["X"], P["s"], P["device"], P["model"], P["world"]
    train_loader, val_loader, test_loader = P["loaders"]
    with _phase("train_model"):
        t_losses, v_losses = _trainer.train_model(
            model, train_loader, val_loader, P["optimizer"], device, num_epochs=P["num_epochs"], is_last=is_last,
            open_browser=open_browser, mode=_cfg.SCATTER_MODE if P["mode"] is None else P["mode"], progress=progress,
            world_size=world)
    with _phase("evaluate_model"):
        test_loss, test_acc = evaluate_model(model, test_loader, device, world_size=world)
        rec_error = compute_reconstruction_error(model, X, s, world_size=world)
    with _phase("compute_alpha_and_norm_ratios"):
        (alpha_val, norm_X_val, norm_ratio_val, rec_scaled, pearson_mean, pearson_std, spearman_mean,
         spearman_std, svd_err, slopes, correlations, spearman_scores, rec_scaled_per_row,
         alpha_per_row) = compute_alpha_and_norm_ratios(model, X, world_size=world)
    rand_indices = P["rand_indices"]
    if rand_indices is None:
        rand_indices = torch.randperm(X.shape[0])[:2]                     # structure.py:390
    sampled_X_rows, sampled_UVT_rows = _metrics.sampled_rows(model, X, rand_indices.tolist())
    gt_loss, gt_acc = compute_ground_truth_metrics(test_loader, X, device, world_size=world)
    return dict((
        ("reconstruction_errors", rec_error), ("log_likelihoods", -test_loss), ("accuracy", test_acc),
        ("gt_log_likelihoods", -gt_loss), ("gt_accuracy", gt_acc), ("train_losses", t_losses),
        ("val_losses", v_losses), ("alpha", alpha_val), ("norm_X", norm_X_val),
        ("norm_ratio", norm_ratio_val), ("reconstruction_error_scaled", rec_scaled),
        ("pearson_corr", pearson_mean), ("pearson_std", pearson_std), ("spearman_corr", spearman_mean),
        ("spearman_std", spearman_std), ("svd_error_scaled", svd_err), ("slopes", slopes),
        ("pearson_corr_matrix", correlations), ("spearman_corr_matrix", spearman_scores),
        ("reconstruction_error_scaled_per_row", rec_scaled_per_row), ("alpha_per_row", alpha_per_row),
        ("sampled_UVT_rows", sampled_UVT_rows), ("sampled_X_rows", sampled_X_rows)))


def run_experiment(n, m, d, p, s, device, lr, weight_decay, reps=5, num_epochs=100, open_browser=False, K=1,
                   d1=None, strategy="random", popularity_method="zipf", alpha=1.5, soft_label=False,
                   generation="base", *, batch_size=64, mode=None, world_size=None):
    """``reps`` independent repetitions of: ground truth -> triplets + BTL labels ->
    train -> evaluate; returns the reference's 23-key dict of per-repetition lists
    (structure.py:420-444).

    Keyword-only extras, reference behaviour by default: ``batch_size`` (the reference
    hard-wires 64, structure.py:668), ``mode`` (scatter mode, None = mfcd_b200.config.SCATTER_MODE), ``world_size``
    (None = torchrun's environment; > 1 = data-parallel: every rank samples and labels the triplets of its own
    user range, the tables are replicated, metrics are combined over the ranks; every rank returns the same dict)."""
    rank, world = _trainer.dist_world(world_size)
    out = {key: [] for key in _RESULT_KEYS}
    for rep in range(reps):
        if d1 is None:
            d1 = d
        P = _prepare_rep(n, m, d, p, s, device, lr, weight_decay, num_epochs, K, strategy, popularity_method, alpha,
                         soft_label, generation, batch_size, mode, rank, world, record=False)
        values = _finish_rep(P, is_last=(rep == reps - 1), open_browser=open_browser)
        for key in _RESULT_KEYS:
            out[key].append(values[key])
    return out


# ---------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------
class BTLPreferenceDataset:
    """(u, i, j, label) samples under the Bradley-Terry-Luce model,
    P(u prefers i over j) = sigmoid(scale * (X[u,i] - X[u,j]))  (structure.py:465-531).
    Hard labels: K consecutive rows per triplet; soft labels (training split
    only): one row with the mean of K Bernoulli draws.  The samples live on the
    GPU as 16-byte records (``.store``); ``.data`` materialises the reference's
    list of tuples on demand."""

    def __init__(self, triplets, X, scale=1.0, K=1, soft_label=False, train=False, *, seed=None):
        self.X = X
        self.scale = scale
        self.soft_label = soft_label
        self.store = self._generate_labels(triplets, K, train=train, seed=seed)
        self._data = None

    def _generate_labels(self, triplets, K, train=False, seed=None):
        soft = bool(self.soft_label and train)
        n, m = self.X.shape
        if _cfg.RNG_MODE == "reference":
            return _reference_labels(triplets, self.X, self.scale, K, soft)
        if not isinstance(triplets, _sampling.TripletSet):
            dev = compute_device(self.X.device if isinstance(self.X, (torch.Tensor, GroundTruth)) else None)
            triplets = _sampling.TripletSet(_sampling.keys_from_triplets(triplets, m, dev), n, m)
        return _sampling.btl_records(self.X, triplets, scale=self.scale, K=K, soft=soft, seed=seed)

    @property
    def data(self):
        if self._data is None:
            u, i, j, z = [c.cpu().tolist() for c in self.store.columns()]
            self._data = list(zip(u, i, j, z))
        return self._data

    def __len__(self):
        return len(self.store)

    def __getitem__(self, idx):
        return self.data[idx]


def _reference_labels(triplets, X, scale, K, soft):
    """Host replay of structure.py:507-519 under the global torch seed: one
    vectorised ``torch.bernoulli`` consumes the generator exactly like the
    reference's per-triplet calls do."""
    trip = triplets.tolist() if isinstance(triplets, _sampling.TripletSet) else list(triplets)
    Xc = X.dense().cpu() if isinstance(X, GroundTruth) else X.detach().cpu()
    if len(trip) == 0:
        e = torch.empty(0, dtype=torch.int64)
        return TripletStore.from_columns(e, e, e, torch.empty(0, dtype=torch.float64))
    idx = torch.tensor(trip, dtype=torch.int64)
    u, i, j = idx[:, 0], idx[:, 1], idx[:, 2]
    score = torch.sigmoid(scale * (Xc[u, i] - Xc[u, j]))
    draws = torch.bernoulli(score.repeat_interleave(K)).view(-1, K)
    if soft:
        z = torch.mean(draws, dim=1).double()
        return TripletStore.from_columns(u, i, j, z)
    return TripletStore.from_columns(u.repeat_interleave(K), i.repeat_interleave(K), j.repeat_interleave(K),
                                     draws.reshape(-1).double())


def get_triplets_from_X(X, num_triplets, strategy="random", exclude=None,
                        popularity_method="zipf", alpha=1.5, n_clusters=10):
    """Strategy dispatch (structure.py:533-588).  Returns a set-like collection of
    unique (u, i, j): a python ``set`` in reference RNG mode, a GPU ``TripletSet``
    otherwise.  Unknown strategies raise ValueError like the reference."""
    exclude = exclude or set()
    if strategy == "random":
        candidates = choose_items_random(X, num_triplets=num_triplets, exclude=exclude)
    elif strategy == "proximity":
        candidates = choose_items_by_proximity(X, num_triplets, exclude)
    elif strategy == "margin":
        candidates = choose_items_by_margin(X, num_triplets, exclude)
    elif strategy == "variance":
        candidates = choose_items_by_variance(X, num_triplets, exclude)
    elif strategy == "popularity":
        candidates = choose_items_by_popularity(X, num_triplets, exclude, method=popularity_method, alpha=alpha)
    elif strategy == "top_k":
        candidates = choose_items_top_k(X, num_triplets, exclude)
    elif strategy == "cluster":
        candidates = choose_items_cluster_based(X, num_triplets, exclude, n_clusters=n_clusters)
    elif strategy == "user_similarity":
        candidates = choose_items_by_user_similarity(X, num_triplets, exclude)
    elif strategy == "svd":
        candidates = choose_items_by_svd_projection(X, num_triplets, exclude)
    else:
        raise ValueError(f"Unknown triplet sampling strategy: {strategy}")
    if isinstance(candidates, _sampling.TripletSet):
        return candidates                     # already unique
    return set(candidates)


# generation name -> (generator function in generation_data, what it returns)
_OTHER_GENERATORS = {
    "low_rank": ("generate_low_rank_matrix", "USV"), "structured": ("generate_structured_embeddings", "UV"),
    "svd": ("generate_svd_embeddings", "UV"), "correlated": ("generate_correlated_embeddings", "UV"),
    "graph": ("generate_graph_embeddings", "UV"), "social": ("generate_social_embeddings", "UV"),
    "temporal": ("generate_temporal_embeddings", "UV"), "hierarchical": ("generate_hierarchical_embeddings", "UV"),
    "gmm": ("generate_gmm_embeddings", "UV"), "clustered": ("generate_clustered_matrix_from_embeddings", "X"),
}


def generate_X(n, m, d, device, generation="base", **kwargs):
    """Ground-truth matrix by scheme (structure.py:590-663).  "base" is the accelerated one; the other ten are
    outside the hot path: they run the reference's host generators when $MFCD_REFERENCE_PATH points at a
    reference checkout and raise NotImplementedError otherwise; unknown names raise ValueError like the reference."""
    if generation == "base":
        return generate_embeddings(n, m, d, device=device)
    if generation in _OTHER_GENERATORS:
        fn_name, returns = _OTHER_GENERATORS[generation]
        fn = getattr(_gen, fn_name)
        if returns == "USV":
            U, V, S = fn(n, m, d, rank=kwargs.get("rank", d), device=device)
            return torch.matmul(torch.matmul(U, torch.diag(S)), V.t())
        if returns == "X":
            return fn(n, m, d, device=device)
        U, V = fn(n, m, d, device=device)
        return torch.matmul(U, V.t())
    raise ValueError(f"Unknown generation method: {generation}")


def _split_permutation(total, device):
    """random_split(..., generator=manual_seed(42)) draws randperm(total) (structure.py:710-713)."""
    if total <= (1 << 24) or _cfg.RNG_MODE == "reference":
        return torch.randperm(total, generator=torch.Generator().manual_seed(42))
    g = torch.Generator(device=device)
    g.manual_seed(42)
    return torch.randperm(total, generator=g, device=device)


def split_dataset_from_triplets(X, num_triplets, scale=1.0, K=1,
                                train_ratio=0.8, val_ratio=0.1,
                                batch_size=64, strategy="random",
                                popularity_method="zipf", alpha=1.5, soft_label=False, *, world_size=1):
    """Sample triplets, split 80/10/10 with the fixed seed 42, top the test split
    up to >= 500 points, label the three splits and wrap them in loaders
    (structure.py:666-742).  Returns (train_loader, val_loader, test_loader).

    world_size > 1 (data parallel): this rank runs the same pipeline on ITS slice of the
    users -- rows [lo, hi) of X, num_triplets / world of the triplets -- so the ranks hold disjoint shards (a
    triplet's key contains its user) and no cross-rank dedup is needed.  ``batch_size`` stays the GLOBAL batch."""
    rank, world = _trainer.dist_world(world_size)
    if world > 1:
        lo, hi = _mdist.split_even(X.shape[0], world, rank)
        a, b = _mdist.split_even(num_triplets, world, rank)
        Xs = X.row_slice(lo, hi) if isinstance(X, GroundTruth) else X[lo:hi]
        loaders = split_dataset_from_triplets(Xs, b - a, scale=scale, K=K, train_ratio=train_ratio,
                                              val_ratio=val_ratio, batch_size=batch_size, strategy=strategy,
                                              popularity_method=popularity_method, alpha=alpha,
                                              soft_label=soft_label, world_size=1)
        for ld in loaders:
            ld.store.rec[:, 0] += lo                      # local user ids -> global
        return loaders
    n, m = X.shape
    found = get_triplets_from_X(X, num_triplets, strategy=strategy, popularity_method=popularity_method,
                                alpha=alpha)
    on_gpu = isinstance(found, _sampling.TripletSet)
    triplets = found if on_gpu else list(found)
    if len(triplets) < num_triplets:
        print(f"⚠️ Only {len(triplets)} triplets generated for strategy: {strategy} (target={num_triplets})")

    total = len(triplets)
    train_size = int(train_ratio * total)
    val_size = int(val_ratio * total)
    if on_gpu:
        perm = _split_permutation(total, triplets.keys.device).to(triplets.keys.device)
        train_triplets = triplets[perm[:train_size]]
        val_triplets = triplets[perm[train_size:train_size + val_size]]
        test_triplets = triplets[perm[train_size + val_size:]]
    else:
        perm = _split_permutation(total, None).tolist()
        train_triplets = [triplets[k] for k in perm[:train_size]]
        val_triplets = [triplets[k] for k in perm[train_size:train_size + val_size]]
        test_triplets = [triplets[k] for k in perm[train_size + val_size:]]

    MIN_TEST_POINTS = 500
    if len(test_triplets) * K < MIN_TEST_POINTS:
        needed = (MIN_TEST_POINTS + K - 1) // K - len(test_triplets)
        if on_gpu:
            seen = _sampling.TripletSet(triplets.keys, n, m)          # train + val + test == everything sampled
            extra = get_triplets_from_X(X, needed, strategy=strategy, popularity_method=popularity_method,
                                        alpha=alpha, exclude=seen)
            test_triplets = _sampling.TripletSet(torch.cat([test_triplets.keys, extra.keys]), n, m)
        else:
            seen = set(train_triplets + val_triplets + test_triplets)
            extra = get_triplets_from_X(X, needed, strategy=strategy, popularity_method=popularity_method,
                                        alpha=alpha, exclude=seen)
            test_triplets += list(extra)

    train_dataset = BTLPreferenceDataset(train_triplets, X, scale=scale, K=K, soft_label=soft_label, train=True)
    val_dataset = BTLPreferenceDataset(val_triplets, X, scale=scale, K=K, soft_label=soft_label)
    test_dataset = BTLPreferenceDataset(test_triplets, X, scale=scale, K=K, soft_label=soft_label)

    rng = "reference" if _cfg.RNG_MODE == "reference" else "device"
    train_loader = TripletLoader(train_dataset.store, batch_size=batch_size, shuffle=True, shuffle_rng=rng)
    val_loader = TripletLoader(val_dataset.store, batch_size=batch_size, shuffle=False, shuffle_rng=rng)
    test_loader = TripletLoader(test_dataset.store, batch_size=batch_size, shuffle=False, shuffle_rng=rng)
    return train_loader, val_loader, test_loader


# ---------------------------------------------------------------------------
# training / evaluation (hot path)
# ---------------------------------------------------------------------------
def train_model(model, train_loader, val_loader, optimizer, device, num_epochs=100, is_last=False,
                open_browser=False, *, mode=None, world_size=None):
    """Per-epoch mean-of-batch-means training and validation BCE (structure.py:812-878).
    Keyword-only extras: ``mode`` (scatter mode; None = mfcd_b200.config.SCATTER_MODE, 'auto' = deterministic for
    batches up to 256, atomic above) and ``world_size`` (data parallel; None = torchrun's environment).  The batch
    size is the loader's (``split_dataset_from_triplets(..., batch_size=...)``)."""
    return _trainer.train_model(model, train_loader, val_loader, optimizer, device, num_epochs=num_epochs,
                                is_last=is_last, open_browser=open_browser,
                                mode=_cfg.SCATTER_MODE if mode is None else mode, progress=True,
                                world_size=world_size)


def evaluate_model(model, test_loader, device, *, world_size=None):
    """(test BCE as mean of batch means, accuracy)  (structure.py:881-921)."""
    return _trainer.evaluate_model(model, test_loader, device, world_size=world_size)


def compute_reconstruction_error(model, X, s, *, world_size=None):
    """|| (UV^T - column means) - sX ||_F / || sX ||_F  (structure.py:925-955)."""
    return _metrics.compute_reconstruction_error(model, X, s, world_size=world_size)


def compute_alpha_and_norm_ratios(model, X_init, *, world_size=None):
    """14 alignment metrics between row-centred UV^T and X (structure.py:958-1082)."""
    return _metrics.compute_alpha_and_norm_ratios(model, X_init, world_size=world_size)


def compute_ground_truth_metrics(test_loader, X, device, *, world_size=None):
    """(MSE of sigmoid(X[u,i]-X[u,j]) vs labels, accuracy of the sign)  (structure.py:1085-1127)."""
    return _trainer.compute_ground_truth_metrics(test_loader, X, device, world_size=world_size)


def start_tensorboard(log_dir='runs/matrix_factorization', port=6006, open_browser=True):
    """The reference never calls this (its call sites sit under ``if False:``,
    structure.py:831-834); kept so the name resolves."""
    raise NotImplementedError("TensorBoard launching is outside the hot path")


def evaluate_ground_truth(n, m, p, d, s, device, K, reps=1, strategy="random", popularity_method="zipf",
                          alpha=1.5, soft_label=False, generation="base"):
    """Ground-truth loss / accuracy without training (structure.py:1154-1200)."""
    losses, accuracies = [], []
    for _ in range(reps):
        X = generate_X(n, m, d, device, generation=generation)
        num_triplets = int(n * m * p / 2)
        _, _, test_loader = split_dataset_from_triplets(
            X, num_triplets, scale=s, K=K, strategy=strategy, popularity_method=popularity_method,
            alpha=alpha, soft_label=soft_label)
        gt_loss, gt_acc = compute_ground_truth_metrics(test_loader, X, device)
        losses.append(gt_loss)
        accuracies.append(gt_acc)
    return losses, accuracies


def parameter_scan_ground_truth(n, m, p, d, s, device, K, linear=False, reps=1, strategy="random",
                                popularity_method="zipf", alpha=1.5, soft_label=False, generation="base"):
    """Sweep of evaluate_ground_truth (structure.py:1203-1269); an unsynchronised
    ``linear=True`` request silently falls back to the grid, as in the reference."""
    grid, lists, synchronised = _normalise_grid({
        'n': n, 'm': m, 'p': p, 'd': d, 's': s, 'K': K, 'strategy': strategy,
        'popularity_method': popularity_method, 'alpha': alpha, 'soft_label': soft_label,
        'generation': generation})
    if linear and synchronised:
        configs = [{k: (v[idx] if len(v) > 1 else v[0]) for k, v in grid.items()}
                   for idx in range(len(lists[0]))]
    else:
        configs = [dict(zip(grid.keys(), combo)) for combo in itertools.product(*grid.values())]
    results = []
    for cfg in tqdm(configs, desc="Training Progress"):
        gt_loss, gt_accuracy = evaluate_ground_truth(**cfg, device=device, reps=reps)
        results.append({'params': cfg, 'results': {'gt_loss': gt_loss, 'gt_accuracy': gt_accuracy}})
    return results
