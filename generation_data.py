"""Drop-in for the reference's ``generation_data.py`` -- hot-path subset.

Same function names and signatures as /root/reference/generation_data.py for
the four sampling strategies named by the north star (random :16, margin :46,
popularity :103, svd :131) and the "base" ground-truth generator (:346); the
work runs in the CUDA samplers of mfcd_b200 (kernels K7).  With
``mfcd_b200.config.RNG_MODE == "reference"`` the random / popularity samplers
replay the reference's host generators instead, so a seeded run reproduces the
reference's triplets bit for bit (parity runs at reference scale).

proximity / variance / top_k (the strategies Runs.ipynb cell 18 sweeps) run on the same
GPU machinery (per-user top-k lists or an item law + the shared dedup).  The reference's "not used" strategies
(cluster, user_similarity) and its ten other generators are outside the accelerated path: they are passed through
to the reference's own host code when ``$MFCD_REFERENCE_PATH`` names a reference checkout (so a Runs.ipynb sweep
over ``strategies=[..., "cluster", "svd"]`` runs), and raise NotImplementedError otherwise -- never silently
something else.
"""
import os

os.environ.setdefault("OMP_NUM_THREADS", "4")      # the reference pins this at import (generation_data.py:2)

import math

import numpy as np
import torch

import mfcd_b200  # noqa: F401  (loads libmfcd_b200.so; fails loudly if it is missing)
from mfcd_b200 import config as _cfg
from mfcd_b200 import sampling as _sampling
from mfcd_b200.store import GroundTruth as _GroundTruth


def _as_exclude(exclude):
    return exclude if exclude else None


# === RANDOM ===  (generation_data.py:16-26)
def choose_items_random(X, num_triplets, exclude):
    """Uniform (u, i, j) with i != j, unique, not in ``exclude``."""
    n, m = X.shape
    if _cfg.RNG_MODE == "reference":
        return _sampling.host_sample_random(n, m, num_triplets, exclude or ())
    return _sampling.sample_random(X, num_triplets, _as_exclude(exclude))


# === MARGIN a.k.a. CLOSE-CALL ===  (generation_data.py:46-84)
def choose_items_by_margin(X, num_triplets, exclude, max_attempts=5000_000):
    """Pairs whose scores differ by at most the adaptive margin; gives up after
    ``max_attempts`` candidates like the reference (printed warning, not an error).
    The reference draws from an unseeded ``default_rng()``, so there is no
    stream to replay: both RNG modes use the GPU sampler."""
    return _sampling.sample_margin(X, num_triplets, _as_exclude(exclude), max_attempts=max_attempts)


# === POPULARITY ===  (generation_data.py:103-128)
def choose_items_by_popularity(X, num_triplets, exclude, method="zipf", alpha=1.5):
    """Item pair drawn without replacement from a zipf / exponential / uniform
    law over item index order; user uniform."""
    n, m = X.shape
    if _cfg.RNG_MODE == "reference":
        if method == "zipf":
            probs = 1 / (np.arange(1, m + 1) ** alpha)
        elif method == "exponential":
            probs = np.exp(-alpha * np.arange(m))
        elif method == "uniform":
            probs = np.ones(m)
        else:
            raise ValueError(f"Unknown popularity method: {method}")
        probs = probs / probs.sum()
        items = np.arange(m)
        seen = exclude or ()
        kept = set()
        while len(kept) < num_triplets:          # host replay: same library calls, same streams
            u = torch.randint(0, n, (1,)).item()
            i, j = np.random.choice(items, size=2, replace=False, p=probs).tolist()
            t = (u, i, j)
            if i != j and t not in seen and t not in kept:
                kept.add(t)
        return list(kept)
    _sampling.popularity_cdf(2, method, alpha)     # validates `method` (ValueError like the reference)
    return _sampling.sample_popularity(X, num_triplets, _as_exclude(exclude), method=method, alpha=alpha)


# === SVD ===  (generation_data.py:131-179)
def choose_items_by_svd_projection(X, num_triplets, exclude, rank=10, top_fraction=0.3):
    """Uniform triplets inside the block of the top-30% users x top-30% items by
    latent-projection norm; at most 5*num_triplets attempts."""
    return _sampling.sample_svd(X, num_triplets, _as_exclude(exclude), rank=rank, top_fraction=top_fraction)


_REFERENCE_MODULE = None


def _reference_module():
    """The reference's own generation_data.py, loaded under a private name from $MFCD_REFERENCE_PATH (a checkout
    of MayeulCassier/Matrix-Factorization-With-Comparison-Data), or None.  Only the strategies and generators that
    are outside the accelerated path are taken from it ("pass through (host)")."""
    global _REFERENCE_MODULE
    if _REFERENCE_MODULE is None:
        root = os.environ.get("MFCD_REFERENCE_PATH", "")
        path = os.path.join(root, "generation_data.py") if root else ""
        if not path or not os.path.isfile(path) or os.path.samefile(path, __file__):
            _REFERENCE_MODULE = False
        else:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_mfcd_reference_generation_data", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _REFERENCE_MODULE = mod
    return _REFERENCE_MODULE or None


def _host_matrix(X):
    """what the reference's host code expects: a CPU float tensor"""
    if isinstance(X, _GroundTruth):
        return X.dense().cpu()
    return X.detach().cpu() if isinstance(X, torch.Tensor) else torch.as_tensor(X)


def _outside_hot_path(name, kind="strategy"):
    """Names outside the B200 hot path: run the reference's host implementation when a reference checkout is
    importable ($MFCD_REFERENCE_PATH), else raise NotImplementedError (never silently something else)."""
    def fn(*args, **kwargs):
        ref = _reference_module()
        if ref is None or not hasattr(ref, name):
            raise NotImplementedError(
                f"{name} is outside the B200 hot path (random / margin / popularity / svd / proximity / variance / "
                f"top_k samplers and the 'base' generator are accelerated; see SURVEY.md section 8f). Set "
                f"MFCD_REFERENCE_PATH to a checkout of the reference to pass it through to the host code.")
        if kind == "strategy" and args:
            args = (_host_matrix(args[0]),) + tuple(args[1:])
            if len(args) > 2 and args[2] is not None and not isinstance(args[2], (set, frozenset)):
                args = args[:2] + (set(args[2]),) + args[3:]        # a GPU TripletSet -> python set of tuples
        return getattr(ref, name)(*args, **kwargs)
    fn.__name__ = name
    return fn


# === PROXIMITY a.k.a. MIN-MAX ===  (generation_data.py:29-43)
def choose_items_by_proximity(X, num_triplets, exclude, k=100):
    """i among the user's k highest scores, j among the k lowest."""
    return _sampling.sample_proximity(X, num_triplets, _as_exclude(exclude), k=k)


# === VARIANCE ===  (generation_data.py:87-99)
def choose_items_by_variance(X, num_triplets, exclude):
    """item pair drawn without replacement with probability proportional to the item's variance across users."""
    return _sampling.sample_variance(X, num_triplets, _as_exclude(exclude))


# === TOP-K a.k.a. TOP_10% ===  (generation_data.py:189-224)
def choose_items_top_k(X, num_triplets, exclude, k=None):
    """i != j among the user's top-k items (k = 10% of the items, at least 5); at most 3 x num_triplets attempts."""
    return _sampling.sample_top_k(X, num_triplets, _as_exclude(exclude), k=k)


choose_items_cluster_based = _outside_hot_path("choose_items_cluster_based")
choose_items_by_user_similarity = _outside_hot_path("choose_items_by_user_similarity")


def estimate_k(num_triplets):
    return math.ceil((1 + math.sqrt(1 + 8 * num_triplets)) / 2)


# === ground-truth matrix, "base" scheme ===  (generation_data.py:346-370)
def generate_embeddings(n, m, d, device="cpu"):
    """X = U S V^T * sqrt(n m)/2 with U, V Haar-orthogonal and S = diag(1/sqrt(d)) on
    the first d entries: a rank-d matrix with entries of std ~0.5.

    reference RNG mode: the same scipy ``ortho_group.rvs`` draws as the reference
    (O(n^3); reference-scale sizes only), hence the same X under a numpy seed.
    device mode: only the first d columns of U and V matter, and the first d
    columns of a Haar matrix are the Q factor of an n x d Gaussian -- O(n d^2)
    on the GPU, which is what makes the 10^4..10^5-row configs possible at all.
    Matrices above config.DENSE_X_MAX_ELEMS stay factored (GroundTruth)."""
    if _cfg.RNG_MODE == "reference":
        from scipy.stats import ortho_group
        k = min(n, m)
        sv = np.zeros(k)
        sv[:d] = 1.0 / np.sqrt(d)
        S = np.zeros((n, m))
        S[:k, :k] = np.diag(sv)
        U = ortho_group.rvs(dim=n)
        V = ortho_group.rvs(dim=m)
        X = (U @ S @ V.T) * np.sqrt(n * m) / 2
        Xt = torch.tensor(X, dtype=torch.float32, device=device)
        return _GroundTruth.attach_factors(Xt, torch.tensor(U[:, :d], dtype=torch.float32),
                                           torch.tensor(V[:, :d], dtype=torch.float32),
                                           np.sqrt(n * m) / (2.0 * np.sqrt(d)))
    return generate_low_rank_gpu(n, m, d, device=device)


def generate_low_rank_gpu(n, m, d, device="cpu", seed=None, force_factored=False):
    from mfcd_b200.store import compute_device
    dev = compute_device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(_sampling.fresh_seed() if seed is None else int(seed))
    A, _ = torch.linalg.qr(torch.randn(n, d, generator=g, device=dev, dtype=torch.float32))
    B, _ = torch.linalg.qr(torch.randn(m, d, generator=g, device=dev, dtype=torch.float32))
    scale = math.sqrt(n * m) / (2.0 * math.sqrt(d))
    if force_factored or n * m > _cfg.DENSE_X_MAX_ELEMS:
        return _GroundTruth(A=A, B=B, scale=scale, device=dev)
    X = (A @ B.T) * scale
    X = X.to(device) if torch.device(device).type == "cpu" else X
    return _GroundTruth.attach_factors(X, A, B, scale)


for _name in ("generate_low_rank_matrix", "generate_structured_embeddings", "generate_svd_embeddings",
              "generate_correlated_embeddings", "generate_graph_embeddings", "generate_social_embeddings",
              "generate_temporal_embeddings", "generate_hierarchical_embeddings", "generate_gmm_embeddings",
              "generate_clustered_matrix_from_embeddings"):
    globals()[_name] = _outside_hot_path(_name, kind="generator")
del _name
