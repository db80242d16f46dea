"""Device-resident data for the hot path: ground-truth views, triplet records, loaders.

Replaces the reference's python-side containers:
  * the dense ``X`` tensor handed around by ``run_experiment`` (structure.py:353),
  * ``BTLPreferenceDataset.data`` -- a python list of ``(u, i, j, label)`` tuples
    (structure.py:507-519),
  * the three ``DataLoader`` objects (structure.py:738-740).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import lib, check, ptr, current_stream


def compute_device(device=None) -> torch.device:
    """The CUDA device the kernels run on.  The reference forwards ``device`` to
    ``.to(device)``; here the compute device is always a GPU -- ``'cpu'`` only
    says where user-visible tensors such as X live.  No GPU => loud failure."""
    if not torch.cuda.is_available():
        raise _lib.MfcdError("mfcd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is not None:
        dev = torch.device(device)
        if dev.type == "cuda":
            return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))


class GroundTruth:
    """X as the kernels see it (struct mfcd_xview): dense fp32 matrix, or the
    low-rank product ``scale * A @ B.T`` evaluated on the fly (needed when the
    dense matrix would not be worth materialising, e.g. 100k x 50k = 20 GB)."""

    def __init__(self, X=None, A=None, B=None, scale=1.0, device=None):
        dev = compute_device(device)
        self.device = dev
        if X is not None:
            self.X = X.detach().to(dev, torch.float32).contiguous()
            self.A = self.B = None
            self.shape = tuple(self.X.shape)
        else:
            self.X = None
            self.A = A.detach().to(dev, torch.float32).contiguous()
            self.B = B.detach().to(dev, torch.float32).contiguous()
            assert self.A.shape[1] == self.B.shape[1]
            self.shape = (self.A.shape[0], self.B.shape[0])
        self.scale = float(scale)
        self.factors = None       # dense X that is KNOWN to equal scale * A @ B.T: (A, B, scale), see attach_factors

    @classmethod
    def wrap(cls, X, device=None):
        if isinstance(X, GroundTruth):
            return X
        gt = cls(X=X, device=device)
        tag = getattr(X, "_mfcd_factors", None)
        if tag is not None and tag[3] == X._version:          # untouched since the generator made it
            gt.factors = (tag[0].to(gt.device), tag[1].to(gt.device), float(tag[2]))
        return gt

    @staticmethod
    def attach_factors(X, A, B, scale):
        """Remember, on a dense X produced by a low-rank generator, the factors it was built from
        (X = scale * A @ B.T).  Consumers that only need spectral information (svd_error_scaled, the SVD
        sampler's top sets) then work on d x d cores instead of an O(n m min(n, m)) dense SVD.  The tag is ignored
        as soon as X is modified in place (version counter) and does not survive copies."""
        X._mfcd_factors = (A.detach(), B.detach(), float(scale), X._version)
        return X

    def xview(self) -> _lib.XView:
        v = _lib.XView()
        if self.X is not None:
            v.X, v.ldx, v.A, v.B, v.dx, v.scale = self.X.data_ptr(), self.X.stride(0), None, None, 0, 1.0
        else:
            v.X, v.ldx, v.A, v.B = None, 0, self.A.data_ptr(), self.B.data_ptr()
            v.dx, v.scale = self.A.shape[1], self.scale
        return v

    def rows(self, r0, nr):
        if self.X is not None:
            return self.X[r0:r0 + nr]
        out = torch.empty((nr, self.shape[1]), dtype=torch.float32, device=self.device)
        xv = self.xview()
        check(lib.mfcd_xview_rows(C.byref(xv), r0, nr, self.shape[1], ptr(out), current_stream()), "mfcd_xview_rows")
        return out

    def dense(self):
        return self.X if self.X is not None else self.rows(0, self.shape[0])

    def row_slice(self, lo, hi):
        """Ground truth of users [lo, hi) as a view (no copy): what one data-parallel rank samples from."""
        if self.X is not None:
            return GroundTruth(X=self.X[lo:hi], device=self.device)
        return GroundTruth(A=self.A[lo:hi], B=self.B, scale=self.scale, device=self.device)


class TripletStore:
    """N labelled comparisons as 16-byte records {int32 u, i, j; float z} in HBM."""

    def __init__(self, rec: torch.Tensor):
        assert rec.dtype == torch.int32 and rec.dim() == 2 and rec.shape[1] == 4 and rec.is_cuda
        self.rec = rec.contiguous()

    def __len__(self):
        return self.rec.shape[0]

    @property
    def device(self):
        return self.rec.device

    @classmethod
    def from_columns(cls, u, i, j, z, device=None):
        """int64 u,i,j + float64 z columns (what the reference's collate yields) -> records."""
        dev = compute_device(device)
        cols = [torch.as_tensor(c) for c in (u, i, j)]
        zc = torch.as_tensor(z)
        N = cols[0].numel()
        with torch.cuda.device(dev):
            cu, ci, cj = [c.to(dev, torch.int64, non_blocking=True).contiguous() for c in cols]
            cz = zc.to(dev, torch.float64, non_blocking=True).contiguous()
            rec = torch.empty((N, 4), dtype=torch.int32, device=dev)
            check(lib.mfcd_pack_triplets(ptr(cu), ptr(ci), ptr(cj), ptr(cz), N, ptr(rec), current_stream()),
                  "mfcd_pack_triplets")
        return cls(rec)

    def columns(self):
        """-> (u, i, j) int64 and z float64 device tensors, the reference's batch dtypes."""
        N = len(self)
        dev = self.device
        with torch.cuda.device(dev):
            u = torch.empty(N, dtype=torch.int64, device=dev)
            i = torch.empty_like(u)
            j = torch.empty_like(u)
            z = torch.empty(N, dtype=torch.float64, device=dev)
            check(lib.mfcd_unpack_triplets(ptr(self.rec), N, ptr(u), ptr(i), ptr(j), ptr(z), current_stream()),
                  "mfcd_unpack_triplets")
        return u, i, j, z

    def slice(self, a, b):
        return TripletStore(self.rec[a:b])

    def pack8(self):
        """-> int64 CUDA tensor of 8-byte wire records (hard labels only); raises if a record does not fit."""
        N = len(self)
        out = torch.empty(N, dtype=torch.int64, device=self.device)
        bad = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.mfcd_pack_triplets8(ptr(self.rec), N, ptr(out), ptr(bad), current_stream()), "mfcd_pack_triplets8")
        if int(bad.item()):
            raise _lib.MfcdError("pack8: soft labels or indices beyond 2^23 users / 2^20 items do not fit the 8-byte format")
        return out

    @classmethod
    def from_packed8(cls, packed: torch.Tensor, out_rec: torch.Tensor = None):
        """8-byte wire records (CUDA int64) -> records; `out_rec` reuses an existing (N, 4) int32 buffer."""
        N = packed.numel()
        rec = torch.empty((N, 4), dtype=torch.int32, device=packed.device) if out_rec is None else out_rec
        with torch.cuda.device(packed.device):
            check(lib.mfcd_unpack_triplets8(ptr(packed), N, ptr(rec), current_stream()), "mfcd_unpack_triplets8")
        return cls(rec)

    def pack_wire(self, start, B):
        """Records [start, start+B) (one user-grouped batch, hard labels, items < 65536) -> uint32 CUDA tensor in
        the run-length wire format, trimmed to the words that have to travel (mfcd_pack_wire)."""
        fixed, cap, wsb = C.c_int64(0), C.c_int64(0), C.c_size_t(0)
        check(lib.mfcd_wire_layout(B, C.byref(fixed), C.byref(cap), C.byref(wsb)), "mfcd_wire_layout")
        dev = self.device
        with torch.cuda.device(dev):
            wire = torch.empty(cap.value, dtype=torch.int32, device=dev)
            ws = torch.empty(max(wsb.value, 1), dtype=torch.uint8, device=dev)
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            check(lib.mfcd_pack_wire(ptr(self.rec[start:start + B]), B, ptr(wire), cap.value, ptr(bad), ptr(ws),
                                     wsb.value, current_stream()), "mfcd_pack_wire")
            if int(bad.item()):
                raise _lib.MfcdError("pack_wire: soft labels or item ids >= 65536 do not fit the run-length wire format")
            n_runs = int(wire[0].item())
        return wire[: fixed.value + n_runs]

    @classmethod
    def from_wire(cls, wire: torch.Tensor, B, out_rec: torch.Tensor = None):
        """run-length wire words (CUDA int32/uint32) -> records; `out_rec` reuses an existing (B, 4) int32 buffer."""
        rec = torch.empty((B, 4), dtype=torch.int32, device=wire.device) if out_rec is None else out_rec
        with torch.cuda.device(wire.device):
            check(lib.mfcd_unpack_wire(ptr(wire), B, ptr(rec), current_stream()), "mfcd_unpack_wire")
        return cls(rec)

    def group_by_user(self, batch_size):
        """Reorder the records in place so that inside every batch of `batch_size` consecutive records one
        user's triplets are adjacent (mfcd_group_by_user).  Batch membership is unchanged; K1 then reads
        U[u] and reduces its gradient once per run (`flags` = FLAG_USER_GROUPED).  Only meaningful for
        loaders that walk the store in order (no epoch permutation)."""
        N = len(self)
        batch_size = int(batch_size)
        if N == 0 or getattr(self, "grouped_batch", None) == batch_size:
            return self
        need = C.c_size_t(0)
        check(lib.mfcd_group_by_user_workspace(N, batch_size, C.byref(need)), "mfcd_group_by_user_workspace")
        with torch.cuda.device(self.device):
            ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            check(lib.mfcd_group_by_user(ptr(self.rec), N, batch_size, ptr(ws), need.value, current_stream()),
                  "mfcd_group_by_user")
        self.grouped_batch = batch_size
        self.__dict__.pop("_hot_cache", None)
        return self

    def user_order(self):
        """int64 CUDA index that sorts the records by user (stable); torch.sort = library call, once per dataset."""
        with torch.cuda.device(self.device):
            return torch.sort(self.rec[:, 0], stable=True).indices

    def k1_flags(self, batch_size, perm=None):
        """flags for the atomic-mode K1 on batches of `batch_size` walked in store order"""
        grouped = perm is None and getattr(self, "grouped_batch", None) == int(batch_size)
        return _lib.FLAG_USER_GROUPED if grouped else 0

    def hot_items(self, n_items, d, batch_size, min_hits_per_batch=2048, sample=1 << 24):
        """Item rows that would each receive >= min_hits_per_batch gradient updates per batch
        (popularity-biased sampling): candidates for K1's shared-memory privatisation.
        Returns (item_slot int8[n_items], hot_items int32[H]) or None when the items are spread out.
        Cached per (n_items, d, batch_size)."""
        key = (int(n_items), int(d), int(batch_size))
        cache = self.__dict__.setdefault("_hot_cache", {})
        if key in cache:
            return cache[key]
        cap = C.c_int32(0)
        check(lib.mfcd_max_hot_items(int(d), C.byref(cap)), "mfcd_max_hot_items")
        out = None
        N = len(self)
        if cap.value > 0 and N > 0 and batch_size >= min_hits_per_batch:
            rec = self.rec[: min(N, sample)]
            counts = torch.bincount(rec[:, 1].long(), minlength=n_items) + torch.bincount(rec[:, 2].long(), minlength=n_items)
            per_batch = counts.double() * (float(batch_size) / rec.shape[0])
            k = min(cap.value, n_items)
            top = torch.topk(per_batch, k)
            sel = top.indices[top.values >= min_hits_per_batch]
            if sel.numel() > 0:
                slot = torch.full((n_items,), -1, dtype=torch.int8, device=self.device)
                slot[sel] = torch.arange(sel.numel(), device=self.device).to(torch.int8)
                out = (slot, sel.to(torch.int32).contiguous())
        cache[key] = out
        return out


class _DatasetView:
    """Duck-type of BTLPreferenceDataset for code that pokes at ``loader.dataset``."""

    def __init__(self, store: TripletStore):
        self._store = store
        self._data = None

    def __len__(self):
        return len(self._store)

    @property
    def data(self):
        if self._data is None:
            u, i, j, z = [c.cpu().tolist() for c in self._store.columns()]
            self._data = list(zip(u, i, j, z))
        return self._data

    def __getitem__(self, idx):
        return self.data[idx]


class TripletLoader:
    """Stands in for ``DataLoader(BTLPreferenceDataset, batch_size, shuffle)``
    (structure.py:738-740).  ``len()`` is the number of batches; iterating yields
    ``(u, i, j, z)`` int64/int64/int64/float64 batches like default_collate does,
    so reference-style consumers keep working, but the fast path never iterates:
    it hands ``store.rec`` and an epoch permutation to the C ABI.

    shuffle_rng:
      "reference": the epoch order is drawn exactly like torch's RandomSampler
                   does (a seed from the global CPU RNG, then randperm with that
                   seed), so a seeded run visits the reference's batches;
      "device":    randperm on the GPU (throughput runs).
    """

    def __init__(self, store: TripletStore, batch_size=64, shuffle=False, shuffle_rng="reference",
                 replay_iter_seed=None):
        self.store = store
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.shuffle_rng = shuffle_rng
        # torch's DataLoader draws one int64 "base seed" from the global generator every time an
        # iterator is created (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__), shuffled
        # or not.  Replaying that draw keeps a seeded run on the reference's RNG stream.
        self.replay_iter_seed = (shuffle_rng == "reference") if replay_iter_seed is None else bool(replay_iter_seed)
        self.dataset = _DatasetView(store)

    def __len__(self):
        return (len(self.store) + self.batch_size - 1) // self.batch_size

    def begin_iteration(self):
        """What `iter(DataLoader)` does to the global RNG, then the epoch order:
        returns an int32 device permutation, or None when not shuffling.  Draws recorded ahead of time
        (record_draws) are replayed first."""
        tape = self.__dict__.get("_tape")
        if tape:
            kind, value = tape.popleft()
            assert kind == "iter", "recorded draws are consumed in the order they were recorded"
            return value
        return self._draw_iteration()

    def _draw_iteration(self):
        if self.replay_iter_seed:
            torch.empty((), dtype=torch.int64).random_()                          # _base_seed
        if not self.shuffle:
            return None
        N = len(self.store)
        if self.shuffle_rng == "reference":
            seed = int(torch.empty((), dtype=torch.int64).random_().item())     # RandomSampler.__iter__
            g = torch.Generator()
            g.manual_seed(seed)
            perm = torch.randperm(N, generator=g)
            return perm.to(self.store.device, torch.int32)
        with torch.cuda.device(self.store.device):
            return torch.randperm(N, device=self.store.device, dtype=torch.int32)

    def _draw_epoch_seed(self):
        """device RNG: the seed of an epoch's keyed bijection (one draw from the global CPU generator)"""
        tape = self.__dict__.get("_tape")
        if tape:
            kind, value = tape.popleft()
            assert kind == "seed", "recorded draws are consumed in the order they were recorded"
            return value
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def uses_epoch_records(self, batch_size=None, atomic=True):
        """True when an epoch at this batch size takes the one-pass shuffle + grouping (epoch_records)."""
        B = int(batch_size or self.batch_size)
        N = len(self.store)
        if not atomic or B <= 256 or N == 0 or N >= (1 << 31):
            return False
        if not self.shuffle:
            return True
        cap = C.c_int32(0)
        check(lib.mfcd_epoch_max_batches(C.byref(cap)), "mfcd_epoch_max_batches")
        return (N + B - 1) // B <= cap.value

    def record_draws(self, count=1, batch_size=None, atomic=True):
        """Make, NOW and in order, the generator draws that `count` future epochs / iterations of this loader would
        make, and replay them later: lets a sweep prepare experiments sequentially (so the global RNG streams are
        consumed exactly as in a sequential run) and run their GPU work concurrently."""
        import collections
        tape = self.__dict__.setdefault("_tape", collections.deque())
        seed_kind = (self.shuffle and self.shuffle_rng != "reference"
                     and self.uses_epoch_records(batch_size, atomic))
        for _ in range(count):
            if seed_kind:
                tape.append(("seed", int(torch.randint(0, 2 ** 62, (1,)).item())))
            else:
                tape.append(("iter", self._draw_iteration()))

    def epoch_perm(self):
        return self.begin_iteration()

    # -- throughput batches: shuffle + per-batch user grouping in one pass (csrc/epoch_batches.cu) -----------
    def _user_sorted(self):
        """(records sorted by user, index of each sorted record in the original store or None).
        device RNG: the store order carries no meaning for a shuffled loader, so the store itself is re-ordered
        (no second copy of a multi-GB shard); reference RNG: epoch permutations index the ORIGINAL order, so a
        sorted copy is kept next to it."""
        cached = self.__dict__.get("_sorted")
        if cached is not None and cached[2] is self.store.rec:
            return cached[0], cached[1]
        order = self.store.user_order()
        with torch.cuda.device(self.store.device):
            rec = self.store.rec[order].contiguous()
        if self.shuffle_rng == "reference":
            out = (rec, order)
        else:
            self.store.rec.copy_(rec)
            self.store.__dict__.pop("_hot_cache", None)
            self.dataset._data = None
            out = (self.store.rec, None)
            del rec
        self._sorted = (out[0], out[1], self.store.rec)
        return out

    def epoch_records(self, batch_size=None, seed=None):
        """This epoch's training records laid out batch after batch, each batch grouped by user:
        -> (TripletStore, k1 flags), or None when the one-pass multisplit does not apply (more batches per epoch than
        mfcd_epoch_max_batches; callers then use epoch_perm()).  Consumes the global CPU generator exactly like
        begin_iteration() (reference RNG) or one seed draw (device RNG)."""
        B = int(batch_size or self.batch_size)
        N = len(self.store)
        if N == 0:
            return None
        if not self.shuffle:
            self.begin_iteration()
            return self.store, self.store.k1_flags(B)
        if not self.uses_epoch_records(B):
            return None
        dev = self.store.device
        rec, order = self._user_sorted()
        pos = None
        if self.shuffle_rng == "reference":
            perm = self.begin_iteration()                     # the reference's RandomSampler order
            with torch.cuda.device(dev):
                pos0 = torch.empty(N, dtype=torch.int32, device=dev)
                check(lib.mfcd_invert_perm(ptr(perm), N, ptr(pos0), current_stream()), "mfcd_invert_perm")
                pos = pos0[order].contiguous()                # position of every record of the SORTED copy
            seed = 0
        elif seed is None:
            seed = self._draw_epoch_seed()
        need = C.c_size_t(0)
        check(lib.mfcd_epoch_batches_workspace(N, B, C.byref(need)), "mfcd_epoch_batches_workspace")
        with torch.cuda.device(dev):
            buf = self.__dict__.get("_epoch_buf")
            if buf is None or buf.shape[0] != N or buf.device != dev:
                buf = self._epoch_buf = torch.empty((N, 4), dtype=torch.int32, device=dev)
            ws = self.__dict__.get("_epoch_ws")
            if ws is None or ws.numel() < need.value or ws.device != dev:
                ws = self._epoch_ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=dev)
            check(lib.mfcd_epoch_batches(ptr(rec), N, B, ptr(pos), int(seed), ptr(buf), ptr(ws), ws.numel(),
                                         current_stream()), "mfcd_epoch_batches")
        self.last_epoch_seed = int(seed)
        return TripletStore(buf), _lib.FLAG_USER_GROUPED

    def __iter__(self):
        u, i, j, z = self.store.columns()
        perm = self.epoch_perm()
        if perm is not None:
            p = perm.long()
            u, i, j, z = u[p], i[p], j[p], z[p]
        for s in range(0, len(self.store), self.batch_size):
            e = s + self.batch_size
            yield u[s:e], i[s:e], j[s:e], z[s:e]


class HostTripletLoader:
    """A loader whose records stay in pinned HOST memory and travel to the GPU one batch per optimiser step
    (datasets that live on the host, or do not fit in HBM).  Stands in for ``DataLoader(dataset, batch_size,
    shuffle=False)`` (structure.py:739): batches are visited in the order the host laid them out; shuffling a
    host-resident dataset is the host's job (a reshuffled epoch = a new HostTripletLoader), or upload it once and
    use TripletLoader.  train_model streams it double-buffered: the copy of batch k+1 overlaps the step on batch k.

    fmt (see include/mfcd_b200.h, hostpack.py):
      "records16"  (B, 4) int32 records as they sit in HBM                                16 B / triplet
      "wire8"      uint64 hard-label records, unpacked by mfcd_unpack_triplets8             8 B / triplet
      "wire_rle"   run-length words of a user-grouped batch, decoded by K1 itself        ~4.5 B / triplet
      "wire8_live" the host keeps RAW 16-byte records (pinned or not); every step the batch is packed to the
                   8-byte format by the library's own host threads (mfcd_host_pack_triplets8, csrc/host_pack.cpp)
                   into a ring of pinned staging buffers while the previous batches are in flight   8 B / triplet
                   over PCIe, nothing prepared beforehand.  pack_threads: host threads of the packer (default:
                   the host's cores divided by the ranks on this host, at most 16).  pack_fraction f in (0, 1]:
                   only the first f of every batch is packed, the rest travels as raw records in the same step
                   (needs pinned batches) -- the packer and the PCIe link then work side by side instead of the
                   packer alone setting the pace; "auto" (prepare() measures the packer and the link once and
                   balances them); default 1.0, env MFCD_PACK_FRACTION.
    """

    def __init__(self, batches, sizes, fmt="records16", user_grouped=False, pack_threads=None, pack_fraction=None):
        assert fmt in ("records16", "wire8", "wire_rle", "wire8_live")
        assert len(batches) == len(sizes)
        for b in batches:
            if not (isinstance(b, torch.Tensor) and b.device.type == "cpu" and b.is_contiguous()
                    and (b.is_pinned() or fmt == "wire8_live")):
                raise ValueError("HostTripletLoader needs contiguous PINNED host tensors (torch.Tensor.pin_memory())")
        self.batches, self.sizes, self.fmt = list(batches), [int(x) for x in sizes], fmt
        if pack_threads is None:
            import os
            local = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
            pack_threads = int(os.environ.get("MFCD_PACK_THREADS", "0") or 0) or \
                max(1, min(16, (os.cpu_count() or 1) // max(local, 1)))
        self.pack_threads = int(pack_threads)
        if pack_fraction is None:
            import os
            pack_fraction = os.environ.get("MFCD_PACK_FRACTION", "1.0") or "1.0"
        if isinstance(pack_fraction, str) and pack_fraction != "auto":
            pack_fraction = float(pack_fraction)
        if pack_fraction != "auto" and not (0.0 < float(pack_fraction) <= 1.0):
            raise ValueError("pack_fraction must be in (0, 1] or 'auto'")
        self.pack_fraction = pack_fraction
        if fmt == "wire8_live" and pack_fraction != 1.0 and not all(b.is_pinned() for b in batches):
            raise ValueError("wire8_live with pack_fraction < 1 ships part of every batch raw: the batches must be "
                             "pinned (from_records pins them)")
        self.user_grouped = bool(user_grouped) or fmt == "wire_rle"
        self.batch_size = max(self.sizes) if self.sizes else 0
        self.shuffle = False
        self.dataset = None

    @classmethod
    def from_records(cls, rec, batch_size, fmt="records16", user_grouped=False, pack_threads=None,
                     pack_fraction=None):
        """rec: (N, 4) int32 records (numpy or CPU tensor), cut into batches of batch_size in order.
        wire8 / wire_rle are packed here on the host (hostpack.py); wire_rle groups every batch by user first;
        wire8_live keeps the raw records and packs them step by step while training."""
        import numpy as np
        from . import hostpack
        a = rec.numpy() if isinstance(rec, torch.Tensor) else np.asarray(rec)
        assert a.ndim == 2 and a.shape[1] == 4 and a.dtype == np.int32
        N, B = a.shape[0], int(batch_size)
        sizes = [min(B, N - s0) for s0 in range(0, N, B)]
        if fmt == "wire8_live":
            import os
            pf = pack_fraction if pack_fraction is not None else (os.environ.get("MFCD_PACK_FRACTION", "1.0") or "1.0")
            if pf in (1.0, "1.0", "1"):
                host = rec if isinstance(rec, torch.Tensor) and rec.is_contiguous() else \
                    torch.from_numpy(np.ascontiguousarray(a))
            else:                                   # part of every batch travels raw: pinned, like records16
                host = torch.empty((N, 4), dtype=torch.int32).pin_memory()
                host.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            return cls([host[s0:s0 + B] for s0 in range(0, N, B)], sizes, fmt, user_grouped, pack_threads, pf)
        if fmt == "records16":
            host = torch.empty((N, 4), dtype=torch.int32).pin_memory()
            host.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            return cls([host[s0:s0 + B] for s0 in range(0, N, B)], sizes, fmt, user_grouped)
        if fmt == "wire8":
            # packed once, here, by the library's host threads (csrc/host_pack.cpp; same bits as hostpack.pack8)
            import ctypes as C
            from ._lib import lib, check
            host = torch.empty(N, dtype=torch.int64).pin_memory()
            src = np.ascontiguousarray(a)
            bad = C.c_int32(0)
            check(lib.mfcd_host_pack_triplets8(src.ctypes.data, N, host.data_ptr(), 0, C.byref(bad)),
                  "mfcd_host_pack_triplets8")
            if bad.value:
                raise ValueError("pack8: soft labels or indices beyond 2^23 users / 2^20 items do not fit the 8-byte format")
            return cls([host[s0:s0 + B] for s0 in range(0, N, B)], sizes, fmt, user_grouped)
        words = [hostpack.pack_wire(a[s0:s0 + B] if user_grouped else hostpack.group_by_user(a[s0:s0 + B]))
                 for s0 in range(0, N, B)]
        host = torch.empty(sum(len(w) for w in words), dtype=torch.int32).pin_memory()
        out, o = [], 0
        for w in words:
            host[o:o + len(w)].copy_(torch.from_numpy(w.view(np.int32)))
            out.append(host[o:o + len(w)])
            o += len(w)
        return cls(out, sizes, fmt, True)

    def __len__(self):
        return len(self.batches)

    def n_samples(self):
        return sum(self.sizes)

    def bytes_per_step(self):
        """bytes that cross PCIe per optimiser step (mean over the batches)"""
        if self.fmt == "wire8_live":
            f = self.pack_fraction if self.pack_fraction != "auto" else getattr(self, "_auto_fraction", 1.0)
            return (8.0 * f + 16.0 * (1.0 - f)) * sum(self.sizes) / max(len(self.sizes), 1)
        return sum(b.numel() * b.element_size() for b in self.batches) / max(len(self.batches), 1)

    def begin_iteration(self):
        return None

    def prepare(self, device=None):
        """Allocate the per-loader staging state now (device slots, copy stream, and for wire8_live the ring of
        pinned buffers + the packer thread) instead of at the first epoch.  Optional; returns self."""
        from . import trainer
        trainer.stager_for(self, compute_device(device))
        return self


def as_loader(loader, device=None) -> TripletLoader:
    """Accept a TripletLoader, or a stock torch DataLoader over (u,i,j,z) rows
    (materialised once into records)."""
    if isinstance(loader, (TripletLoader, HostTripletLoader)):
        return loader
    ds = getattr(loader, "dataset", None)
    if ds is None:
        raise TypeError(f"cannot use {type(loader).__name__} as a triplet loader")
    rows = ds.data if hasattr(ds, "data") else [ds[k] for k in range(len(ds))]
    if len(rows) == 0:
        u = i = j = torch.empty(0, dtype=torch.int64)
        z = torch.empty(0, dtype=torch.float64)
    else:
        u = torch.tensor([int(r[0]) for r in rows], dtype=torch.int64)
        i = torch.tensor([int(r[1]) for r in rows], dtype=torch.int64)
        j = torch.tensor([int(r[2]) for r in rows], dtype=torch.int64)
        z = torch.tensor([float(r[3]) for r in rows], dtype=torch.float64)
    store = TripletStore.from_columns(u, i, j, z, device=device)
    sampler = getattr(loader, "sampler", None)
    shuffle = type(sampler).__name__ == "RandomSampler"
    return TripletLoader(store, batch_size=getattr(loader, "batch_size", 64) or 64, shuffle=shuffle)
