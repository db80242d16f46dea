"""GPU triplet samplers (K7), the BTL label sampler (K8) and the host-side
"reference RNG" replay used for seeded parity runs.

Reference: ``choose_items_random`` / ``_by_margin`` / ``_by_popularity`` /
``_by_svd_projection`` (generation_data.py:16-179), ``get_triplets_from_X``
(structure.py:533-588), ``BTLPreferenceDataset`` (structure.py:465-531) and
``split_dataset_from_triplets`` (structure.py:666-742).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, current_stream
from .store import GroundTruth, TripletStore, compute_device

KEY_NONE = -1  # 0xFFFF...F viewed as int64


# ---------------------------------------------------------------------------
# triplet sets
# ---------------------------------------------------------------------------
class TripletSet:
    """A set of unique (u, i, j) triplets held on the GPU as 64-bit keys
    ``(u*m + i)*m + j`` in acceptance order.  Quacks like the python ``set`` /
    ``list`` of tuples the reference passes around (len, iteration, ``in``),
    without ever building tuples unless somebody iterates."""

    def __init__(self, keys: torch.Tensor, n: int, m: int):
        self.keys = keys      # int64 view of the uint64 keys, CUDA
        self.n, self.m = int(n), int(m)

    def __len__(self):
        return int(self.keys.numel())

    def columns(self):
        k = self.keys
        j = k % self.m
        t = k // self.m
        return t // self.m, t % self.m, j

    def tolist(self):
        u, i, j = [c.cpu().tolist() for c in self.columns()]
        return list(zip(u, i, j))

    def __iter__(self):
        return iter(self.tolist())

    def __contains__(self, t):
        u, i, j = t
        key = (int(u) * self.m + int(i)) * self.m + int(j)
        return bool((self.keys == key).any().item())

    def __getitem__(self, idx):
        if isinstance(idx, (slice, torch.Tensor, list, np.ndarray)):
            return TripletSet(self.keys[idx], self.n, self.m)
        return self.tolist()[idx]

    def union_keys(self, other):
        return torch.cat([self.keys, other.keys])


def keys_from_triplets(triplets, m, device):
    """python iterable of (u,i,j) / TripletSet / None -> int64 CUDA key tensor"""
    if triplets is None:
        return torch.empty(0, dtype=torch.int64, device=device)
    if isinstance(triplets, TripletSet):
        return triplets.keys.to(device)
    if len(triplets) == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    arr = np.asarray(list(triplets), dtype=np.int64).reshape(-1, 3)
    k = (arr[:, 0] * m + arr[:, 1]) * m + arr[:, 2]
    return torch.from_numpy(k).to(device)


# ---------------------------------------------------------------------------
# K7 driver: rounds of candidates + sequential-accept dedup
# ---------------------------------------------------------------------------
_seed_counter = [0]


def fresh_seed():
    """A 63-bit seed drawn from torch's global CPU generator: seeding torch
    makes the GPU samplers reproducible too, like the reference's host samplers."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _accept_rounds(draw, n, m, num_triplets, exclude_keys, device, max_attempts=None, first_rate=1.0):
    """draw(count, counter0, out_keys) fills candidate keys; returns accepted keys."""
    have = torch.empty(0, dtype=torch.int64, device=device)
    if num_triplets <= 0:
        return have, 0
    seen_base = exclude_keys
    attempts = 0
    rate = max(min(first_rate, 1.0), 1e-4)
    n_out = torch.zeros(1, dtype=torch.int64, device=device)
    stall = 0
    while have.numel() < num_triplets:
        need = num_triplets - have.numel()
        count = int(min(max(1024, math.ceil(need / rate * 1.15) + 64), 1 << 27))
        if max_attempts is not None:
            count = min(count, max_attempts - attempts)
            if count <= 0:
                break
        cand = torch.empty(count, dtype=torch.int64, device=device)
        draw(count, attempts, cand)
        seen = torch.cat([seen_base, have]) if (seen_base.numel() or have.numel()) else seen_base
        ws_bytes = C.c_size_t(0)
        check(lib.mfcd_unique_workspace_bytes(seen.numel(), count, C.byref(ws_bytes)), "mfcd_unique_workspace_bytes")
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=device)
        out = torch.empty(need, dtype=torch.int64, device=device)
        check(lib.mfcd_unique_accept(ptr(seen) if seen.numel() else None, seen.numel(), ptr(cand), count, need,
                                     ptr(out), ptr(n_out), ptr(ws), ws.numel(), current_stream()),
              "mfcd_unique_accept")
        got = int(n_out.item())
        attempts += count
        have = torch.cat([have, out[:got]])
        rate = max(got / count, 1e-4) if got else max(rate * 0.25, 1e-4)
        stall = stall + 1 if got == 0 else 0
        if stall >= 8:        # the strategy cannot deliver more (saturated block / margin)
            break
    return have, attempts


def sample_random(X, num_triplets, exclude=None, seed=None):
    gt = GroundTruth.wrap(X)
    n, m = gt.shape
    dev = gt.device
    seed = fresh_seed() if seed is None else seed
    with torch.cuda.device(dev):
        def draw(count, c0, out):
            check(lib.mfcd_sample_random(n, m, count, seed, c0, ptr(out), current_stream()), "mfcd_sample_random")
        keys, _ = _accept_rounds(draw, n, m, num_triplets, keys_from_triplets(exclude, m, dev), dev)
    return TripletSet(keys, n, m)


def margin_threshold(gt: GroundTruth, num_triplets):
    """mean over the first min(10,n) rows of (max - min), times N/(n m)  (generation_data.py:56-57)."""
    n, m = gt.shape
    sample = gt.rows(0, min(10, n))
    spread = (sample.max(dim=1).values - sample.min(dim=1).values).mean().item()
    return float(np.float32(spread) * num_triplets / (n * m))


def sample_margin(X, num_triplets, exclude=None, max_attempts=5_000_000, seed=None):
    gt = GroundTruth.wrap(X)
    n, m = gt.shape
    dev = gt.device
    seed = fresh_seed() if seed is None else seed
    margin = margin_threshold(gt, num_triplets)
    xv = gt.xview()
    with torch.cuda.device(dev):
        def draw(count, c0, out):
            check(lib.mfcd_sample_margin(n, m, count, seed, c0, C.byref(xv), margin, ptr(out), current_stream()),
                  "mfcd_sample_margin")
        keys, attempts = _accept_rounds(draw, n, m, num_triplets, keys_from_triplets(exclude, m, dev), dev,
                                        max_attempts=max_attempts, first_rate=0.25)
    if keys.numel() < num_triplets:
        print(f"⚠️ Only {keys.numel()} triplets generated (target={num_triplets}, margin={margin:.4f}) "
              f"after {attempts} attempts.")
    return TripletSet(keys, n, m)


def popularity_cdf(m, method="zipf", alpha=1.5):
    """float64 probabilities over item index order (generation_data.py:110-119), as inclusive prefix sums."""
    if method == "zipf":
        probs = 1.0 / (np.arange(1, m + 1) ** alpha)
    elif method == "exponential":
        probs = np.exp(-alpha * np.arange(m))
    elif method == "uniform":
        probs = np.ones(m)
    else:
        raise ValueError(f"Unknown popularity method: {method}")
    probs = probs / probs.sum()
    return np.cumsum(probs)


def sample_popularity(X, num_triplets, exclude=None, method="zipf", alpha=1.5, seed=None, max_attempts=None):
    gt = GroundTruth.wrap(X)
    n, m = gt.shape
    dev = gt.device
    seed = fresh_seed() if seed is None else seed
    cdf = torch.from_numpy(popularity_cdf(m, method, alpha)).to(dev)
    with torch.cuda.device(dev):
        def draw(count, c0, out):
            check(lib.mfcd_sample_popularity(n, m, count, seed, c0, ptr(cdf), ptr(out), current_stream()),
                  "mfcd_sample_popularity")
        keys, _ = _accept_rounds(draw, n, m, num_triplets, keys_from_triplets(exclude, m, dev), dev,
                                 max_attempts=max_attempts, first_rate=0.5)
    return TripletSet(keys, n, m)


def svd_top_sets(gt: GroundTruth, rank, top_fraction=0.3):
    """Top users / items by the row norms of U_k S_k and V_k S_k (generation_data.py:149-162).
    The truncated SVD is a library call (the reference uses ARPACK svds)."""
    n, m = gt.shape
    if gt.X is None or gt.factors is not None:
        A, B, scale = (gt.A, gt.B, gt.scale) if gt.X is None else gt.factors
        qa, ra = torch.linalg.qr(A.double())
        qb, rb = torch.linalg.qr(B.double())
        uc, s, vch = torch.linalg.svd(scale * (ra @ rb.T))
        k = min(rank, s.numel())
        user_norms = torch.linalg.norm((qa @ uc[:, :k]) * s[:k], dim=1)
        item_norms = torch.linalg.norm((qb @ vch.T[:, :k]) * s[:k], dim=1)
    else:
        Uf, S, Vh = torch.linalg.svd(gt.X.double(), full_matrices=False)
        user_norms = torch.linalg.norm(Uf[:, :rank] * S[:rank], dim=1)
        item_norms = torch.linalg.norm(Vh[:rank].T * S[:rank], dim=1)
    nu = max(1, int(top_fraction * n))
    ni = max(2, int(top_fraction * m))
    top_users = torch.argsort(user_norms)[-nu:].to(torch.int32).contiguous()
    top_items = torch.argsort(item_norms)[-ni:].to(torch.int32).contiguous()
    return top_users, top_items


def sample_svd(X, num_triplets, exclude=None, rank=10, top_fraction=0.3, seed=None):
    gt = GroundTruth.wrap(X)
    n, m = gt.shape
    dev = gt.device
    rank = int(num_triplets / (n * m) * max(n, m))            # the argument is overridden (generation_data.py:144)
    if not (0 < rank < min(n, m)):
        raise ValueError(f"`k` must be an integer satisfying `0 < k < min(A.shape)`. (k={rank})")   # svds' error
    seed = fresh_seed() if seed is None else seed
    top_users, top_items = svd_top_sets(gt, rank, top_fraction)
    with torch.cuda.device(dev):
        def draw(count, c0, out):
            check(lib.mfcd_sample_block(m, count, seed, c0, ptr(top_users), top_users.numel(), ptr(top_items),
                                        top_items.numel(), ptr(out), current_stream()), "mfcd_sample_block")
        keys, _ = _accept_rounds(draw, n, m, num_triplets, keys_from_triplets(exclude, m, dev), dev,
                                 max_attempts=num_triplets * 5)
    if keys.numel() < num_triplets:
        print(f"⚠️ Only {keys.numel()} triplets generated (target={num_triplets})")
    return TripletSet(keys, n, m)


def _topk_lists(gt: GroundTruth, k, largest=True, chunk_rows=4096):
    """(n x k) int32 table of each user's k largest (or smallest) items by X[u] (torch.topk: library call)."""
    n, m = gt.shape
    out = torch.empty((n, k), dtype=torch.int32, device=gt.device)
    for r0 in range(0, n, chunk_rows):
        rows = gt.rows(r0, min(chunk_rows, n - r0))
        out[r0:r0 + rows.shape[0]] = torch.topk(rows, k, dim=1, largest=largest).indices.to(torch.int32)
    return out.contiguous()


def _sample_lists(gt, num_triplets, exclude, list_i, list_j, same_list, seed, max_attempts=None):
    n, m = gt.shape
    dev = gt.device
    seed = fresh_seed() if seed is None else seed
    with torch.cuda.device(dev):
        def draw(count, c0, out):
            check(lib.mfcd_sample_lists(n, m, count, seed, c0, ptr(list_i), list_i.shape[1], ptr(list_j),
                                        list_j.shape[1], int(same_list), ptr(out), current_stream()), "mfcd_sample_lists")
        keys, attempts = _accept_rounds(draw, n, m, num_triplets, keys_from_triplets(exclude, m, dev), dev,
                                        max_attempts=max_attempts)
    return TripletSet(keys, n, m), attempts


def sample_proximity(X, num_triplets, exclude=None, k=100, seed=None):
    """i among the user's top-k items, j among the bottom-k (generation_data.py:29-43)."""
    gt = GroundTruth.wrap(X)
    kk = min(k, gt.shape[1])
    ts, _ = _sample_lists(gt, num_triplets, exclude, _topk_lists(gt, kk, True), _topk_lists(gt, kk, False), False, seed)
    return ts


def sample_top_k(X, num_triplets, exclude=None, k=None, seed=None):
    """i != j among the user's top-k items, k = min(m, max(5, int(0.1 m))), at most 3 x num_triplets attempts
    (generation_data.py:189-224)."""
    gt = GroundTruth.wrap(X)
    m = gt.shape[1]
    if k is None:
        k = min(m, max(5, int(0.1 * m)))
    top = _topk_lists(gt, k, True)
    ts, _ = _sample_lists(gt, num_triplets, exclude, top, top, True, seed, max_attempts=num_triplets * 3)
    if len(ts) < num_triplets:
        print(f"⚠️ Only {len(ts)} triplets generated (target={num_triplets}, k={k})")
    return ts


def sample_variance(X, num_triplets, exclude=None, seed=None):
    """item pair without replacement from probabilities proportional to the item's variance across users
    (generation_data.py:87-99): the popularity sampler with a different law."""
    gt = GroundTruth.wrap(X)
    n, m = gt.shape
    dev = gt.device
    var = torch.var(gt.dense(), dim=0).double()
    cdf = torch.cumsum(var / var.sum(), dim=0).contiguous()
    seed = fresh_seed() if seed is None else seed
    with torch.cuda.device(dev):
        def draw(count, c0, out):
            check(lib.mfcd_sample_popularity(n, m, count, seed, c0, ptr(cdf), ptr(out), current_stream()),
                  "mfcd_sample_popularity")
        keys, _ = _accept_rounds(draw, n, m, num_triplets, keys_from_triplets(exclude, m, dev), dev, first_rate=0.5)
    return TripletSet(keys, n, m)


# ---------------------------------------------------------------------------
# K8: labels
# ---------------------------------------------------------------------------
def btl_records(X, triplets: TripletSet, scale=1.0, K=1, soft=False, seed=None, uniforms=None) -> TripletStore:
    """Labelled records for a TripletSet: K consecutive hard labels per triplet, or
    one soft label (mean of K draws) -- structure.py:507-519."""
    gt = GroundTruth.wrap(X)
    dev = gt.device
    N = len(triplets)
    n_out = N if soft else N * K
    rec = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
    seed = fresh_seed() if seed is None else seed
    xv = gt.xview()
    keys = triplets.keys.to(dev).contiguous()
    if uniforms is not None:
        uniforms = torch.as_tensor(uniforms).to(dev, torch.float32).contiguous()
        assert uniforms.numel() == N * K
    with torch.cuda.device(dev):
        check(lib.mfcd_btl_labels(C.byref(xv), ptr(keys), N, gt.shape[1], int(K), float(scale), int(bool(soft)),
                                  seed, ptr(uniforms), ptr(rec), current_stream()), "mfcd_btl_labels")
    return TripletStore(rec)


# ---------------------------------------------------------------------------
# "reference RNG": replay of torch's CPU Mersenne Twister, vectorised
# ---------------------------------------------------------------------------
class TorchMT:
    """Reads raw 32-bit outputs from torch's global CPU generator in bulk and
    leaves the generator exactly where a one-at-a-time python loop would have.

    torch.randint(0, r, ...) with r < 2^28... consumes one 32-bit word per element
    and returns ``word % r``; numpy's MT19937 tempering is the same function, so
    the state is moved into numpy, advanced there, and written back."""

    _SEED_OFF, _LEFT_OFF, _SEEDED_OFF, _NEXT_OFF, _STATE_OFF = 0, 8, 12, 16, 24

    def __init__(self):
        st = torch.get_rng_state().numpy().copy()
        self._raw = st
        left = int(st[self._LEFT_OFF:self._LEFT_OFF + 4].view(np.int32)[0])
        key = st[self._STATE_OFF:self._STATE_OFF + 624 * 8].view(np.uint64).astype(np.uint32)
        self._bg = np.random.MT19937()
        self._bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 625 - left}}

    def words(self, count):
        return self._bg.random_raw(int(count)).astype(np.uint64)

    def commit(self, consumed_back=0):
        """write the advanced state back into torch (optionally un-reading words)"""
        s = self._bg.state["state"]
        pos = int(s["pos"])
        st = self._raw
        st[self._STATE_OFF:self._STATE_OFF + 624 * 8] = s["key"].astype(np.uint64).view(np.uint8)
        st[self._LEFT_OFF:self._LEFT_OFF + 4] = np.array([625 - pos], np.int32).view(np.uint8)
        st[self._NEXT_OFF:self._NEXT_OFF + 8] = np.array([pos], np.uint64).view(np.uint8)
        torch.set_rng_state(torch.from_numpy(st))


def host_sample_random(n, m, num_triplets, exclude=()):
    """Bit-exact replay of choose_items_random (generation_data.py:16-26) under
    the current torch seed: attempt k consumes three generator words
    (u = w0 % n ; i = w1 % m ; j = w2 % m) and is kept iff i != j and new.
    Returns the kept triplets in the iteration order of the reference's
    ``list(set)``."""
    seen = set(exclude)
    kept = set()
    block = max(256, int(num_triplets * 1.2))
    while len(kept) < num_triplets:
        mt = TorchMT()
        w = mt.words(3 * block).reshape(block, 3)
        us, is_, js = (w[:, 0] % n).tolist(), (w[:, 1] % m).tolist(), (w[:, 2] % m).tolist()
        used = 0
        for u, i, j in zip(us, is_, js):
            used += 1
            t = (u, i, j)
            if i != j and t not in seen and t not in kept:
                kept.add(t)
                if len(kept) >= num_triplets:
                    break
        # rewind: re-read the state and consume exactly 3*used words
        mt2 = TorchMT()
        mt2.words(3 * used)
        mt2.commit()
    return list(kept)
