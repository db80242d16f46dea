"""Model, optimiser glue and the training / evaluation drivers of the hot path.

Mirrors (same names, arguments and return values) the reference's
``MatrixFactorization`` (structure.py:746-795), ``train_model`` (:812-878),
``evaluate_model`` (:881-921) and ``compute_ground_truth_metrics`` (:1085-1127);
the arithmetic runs in the CUDA kernels behind include/mfcd_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, current_stream, wait_for_stream
from .store import GroundTruth, HostTripletLoader, TripletLoader, TripletStore, as_loader, compute_device

MODE_ATOMIC = 0
MODE_DETERMINISTIC = 1


def resolve_mode(mode, batch_size):
    """'atomic' | 'deterministic' | 'auto' (env MFCD_MODE overrides 'auto').
    auto: deterministic for the reference's small batches (bit-reproducible like
    the CPU reference), atomic red.global scatter for throughput-sized batches."""
    if mode in (None, "auto"):
        mode = os.environ.get("MFCD_MODE", "auto")
    if mode == "auto":
        return MODE_DETERMINISTIC if batch_size <= 256 else MODE_ATOMIC
    if mode in ("atomic", MODE_ATOMIC):
        return MODE_ATOMIC
    if mode in ("deterministic", "det", MODE_DETERMINISTIC):
        return MODE_DETERMINISTIC
    raise ValueError(f"unknown scatter mode: {mode!r}")


class MatrixFactorization(nn.Module):
    """sigma(<U_u, V_i - V_j>) with U in R^{n x d}, V in R^{m x d}  (structure.py:746-795).

    Parameters are drawn exactly like the reference does (``randn(n,d)/sqrt(d)``
    from the global CPU generator, U first), so a seeded run starts from the same
    weights.  For training the two tables are packed into one flat CUDA buffer
    (U then V) of which ``self.U`` / ``self.V`` are views: one optimiser launch
    and one gradient all-reduce cover both.
    """

    def __init__(self, n_users, n_items, d):
        super().__init__()
        scale = torch.sqrt(torch.tensor(d, dtype=torch.float32))
        self.U = nn.Parameter(torch.randn(n_users, d) / scale)
        self.V = nn.Parameter(torch.randn(n_items, d) / scale)
        self._flat = None

    # -- flat CUDA storage ---------------------------------------------------
    def flat_state(self, device=None, storage=None):
        """storage = (params, grads): externally allocated flat fp32 CUDA buffers to live in."""
        n, d = self.U.shape
        m = self.V.shape[0]
        fs = self._flat
        if (storage is None and fs is not None and fs.params.is_cuda and self.U.data_ptr() == fs.params.data_ptr()
                and self.V.data_ptr() == fs.params.data_ptr() + 4 * n * d):
            return fs
        if storage is None and fs is not None and fs.host_mirror is not None and fs.host_mirror == self._mirror_key():
            return fs           # U / V are the untouched host copies mirror_to_host() left: the flat state is current
        dev = compute_device(device if device is not None else (self.U.device if self.U.is_cuda else None))
        fs = _FlatState(n, m, d, dev, *(storage or ()))
        with torch.no_grad():
            fs.params[: n * d].view(n, d).copy_(self.U.detach())
            fs.params[n * d:].view(m, d).copy_(self.V.detach())
        self.U.data = fs.params[: n * d].view(n, d)
        self.V.data = fs.params[n * d:].view(m, d)
        self._flat = fs
        return fs

    def _mirror_key(self):
        return (self.U.data_ptr(), self.U._version, self.V.data_ptr(), self.V._version)

    def mirror_to_host(self):
        """``device='cpu'`` callers get CPU ``model.U`` / ``model.V`` back, like the reference's
        ``model.to('cpu')`` would leave them (reference-style post-processing against a CPU X keeps working);
        the flat CUDA state stays attached and is reused as long as the host copies are not modified."""
        fs = self._flat
        if fs is None or not fs.params.is_cuda:
            return
        n, d, m = fs.n, fs.d, fs.m
        with torch.no_grad():
            self.U.data = fs.params[: n * d].view(n, d).cpu()
            self.V.data = fs.params[n * d:].view(m, d).cpu()
        fs.host_mirror = self._mirror_key()

    def forward(self, u, i, j):
        """Preference probabilities for index tensors u, i, j (inference only: the
        training path is the fused forward/backward kernel, not autograd)."""
        fs = self.flat_state()
        dev = fs.params.device
        u = torch.as_tensor(u).to(dev, torch.int64).contiguous()
        i = torch.as_tensor(i).to(dev, torch.int64).contiguous()
        j = torch.as_tensor(j).to(dev, torch.int64).contiguous()
        out = torch.empty(u.numel(), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mfcd_triplet_scores(ptr(fs.U), ptr(fs.V), ptr(u), ptr(i), ptr(j), u.numel(), fs.d, ptr(out),
                                          current_stream()), "mfcd_triplet_scores")
        return out


class _FlatState:
    """(n+m)*d fp32 parameters, gradients and two optimiser moments, U first."""

    def __init__(self, n, m, d, device, params=None, grads=None):
        self.n, self.m, self.d = n, m, d
        numel = (n + m) * d
        # params / grads may be handed in (symmetric-memory buffers shared with peer GPUs, dist.PeerExchange)
        self.params = torch.zeros(numel, dtype=torch.float32, device=device) if params is None else params[:numel]
        self.grads = torch.zeros(numel, dtype=torch.float32, device=device) if grads is None else grads[:numel]
        self.state1 = torch.zeros(numel, dtype=torch.float32, device=device)
        self.state2 = torch.zeros(numel, dtype=torch.float32, device=device)
        self.step = 0
        self.workspace = None
        self.host_mirror = None

    @property
    def U(self):
        return self.params[: self.n * self.d]

    @property
    def V(self):
        return self.params[self.n * self.d:]

    def ensure_workspace(self, nbytes):
        if nbytes and (self.workspace is None or self.workspace.numel() < nbytes):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.params.device)
        return self.workspace


class OptimizerSpec:
    """Hyper-parameters read off a torch.optim.Adam / SGD instance (the object
    run_experiment builds at structure.py:364)."""

    def __init__(self, optimizer):
        if isinstance(optimizer, OptimizerSpec):
            self.__dict__.update(optimizer.__dict__)
            return
        if len(optimizer.param_groups) != 1:
            raise NotImplementedError("one param group (the reference's setup) is supported")
        g = optimizer.param_groups[0]
        name = type(optimizer).__name__
        self.momentum = 0.0
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-8
        if name == "Adam":
            if g.get("amsgrad", False) or g.get("maximize", False) or g.get("decoupled_weight_decay", False):
                raise NotImplementedError("amsgrad / maximize / decoupled weight decay are not implemented")
            self.kind = 0
            self.beta1, self.beta2 = g["betas"]
            self.eps = g["eps"]
        elif name == "SGD":
            if g.get("nesterov", False) or g.get("dampening", 0) != 0 or g.get("maximize", False):
                raise NotImplementedError("nesterov / dampening / maximize are not implemented")
            self.kind = 1
            self.momentum = g.get("momentum", 0.0)
        else:
            raise NotImplementedError(f"optimizer {name} has no fused kernel (Adam and SGD do)")
        self.lr = float(g["lr"])
        self.weight_decay = float(g.get("weight_decay", 0.0))

    @classmethod
    def adam(cls, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        s = cls.__new__(cls)
        s.kind, s.lr, s.weight_decay, s.momentum = 0, float(lr), float(weight_decay), 0.0
        s.beta1, s.beta2, s.eps = betas[0], betas[1], eps
        return s


def _import_optimizer_state(model, optimizer, fs):
    """Continue from a torch optimiser that already holds state for U / V."""
    if isinstance(optimizer, OptimizerSpec) or not hasattr(optimizer, "state"):
        return
    n, d, m = fs.n, fs.d, fs.m
    if not optimizer.state.get(model.U, None) and not optimizer.state.get(model.V, None):
        # a fresh torch optimiser starts from zero moments and step 0 (the reference builds a new Adam per
        # repetition, structure.py:364) -- do not carry over what an earlier optimiser left in the flat state
        fs.state1.zero_()
        fs.state2.zero_()
        fs.step = 0
        return
    for p, sl in ((model.U, slice(0, n * d)), (model.V, slice(n * d, (n + m) * d))):
        st = optimizer.state.get(p, None)
        if not st:
            continue
        if "exp_avg" in st and st["exp_avg"].data_ptr() != fs.state1[sl].data_ptr():
            fs.state1[sl].copy_(st["exp_avg"].reshape(-1))
            fs.state2[sl].copy_(st["exp_avg_sq"].reshape(-1))
        if "momentum_buffer" in st and st["momentum_buffer"] is not None \
                and st["momentum_buffer"].data_ptr() != fs.state1[sl].data_ptr():
            fs.state1[sl].copy_(st["momentum_buffer"].reshape(-1))
        if "step" in st:
            fs.step = max(fs.step, int(float(st["step"])))


def _export_optimizer_state(model, optimizer, fs, spec):
    """Leave the torch optimiser object consistent with what the kernels did."""
    if isinstance(optimizer, OptimizerSpec) or not hasattr(optimizer, "state"):
        return
    n, d, m = fs.n, fs.d, fs.m
    for p, sl, shape in ((model.U, slice(0, n * d), (n, d)), (model.V, slice(n * d, (n + m) * d), (m, d))):
        if spec.kind == 0:
            optimizer.state[p] = {"step": torch.tensor(float(fs.step)),
                                  "exp_avg": fs.state1[sl].view(shape),
                                  "exp_avg_sq": fs.state2[sl].view(shape)}
        elif spec.momentum != 0.0:
            optimizer.state[p] = {"momentum_buffer": fs.state1[sl].view(shape)}


def run_epoch(fs: _FlatState, store: TripletStore, perm, batch_size, spec: OptimizerSpec, mode, *, flags=None,
              hot="auto"):
    """One epoch of structure.py:845-852 through mfcd_train_epoch; returns the
    per-step batch-mean losses as a float32 device tensor.
    flags: MFCD_FLAG_* for the atomic K1 (None = what the store knows about its own layout);
    hot:   (item_slot, hot_items) for hot-row privatisation, None = off, "auto" = counted on `store`."""
    N = len(store)
    n_steps = (N + batch_size - 1) // batch_size
    dev = fs.params.device
    losses = torch.zeros(max(n_steps, 1), dtype=torch.float32, device=dev)
    if n_steps == 0:
        return losses[:0]
    a = _lib.EpochArgs()
    a.params, a.grads, a.state1, a.state2 = ptr(fs.params), ptr(fs.grads), ptr(fs.state1), ptr(fs.state2)
    a.n_users, a.n_items, a.d = fs.n, fs.m, fs.d
    a.optimizer, a.mode = spec.kind, mode
    a.flags = (store.k1_flags(batch_size, perm) if flags is None else int(flags)) if mode == MODE_ATOMIC else 0
    a.rec, a.perm = ptr(store.rec), ptr(perm)
    a.n_samples, a.batch_size = N, batch_size
    a.lr, a.beta1, a.beta2, a.eps = spec.lr, spec.beta1, spec.beta2, spec.eps
    a.weight_decay, a.momentum = spec.weight_decay, spec.momentum
    a.step0 = fs.step
    a.step_losses = ptr(losses)
    ws_bytes = C.c_size_t(0)
    check(lib.mfcd_train_epoch_workspace(C.byref(a), C.byref(ws_bytes)), "mfcd_train_epoch_workspace")
    ws = fs.ensure_workspace(ws_bytes.value)
    a.workspace, a.workspace_bytes = ptr(ws), (ws.numel() if ws is not None else 0)
    if mode != MODE_ATOMIC:
        hot = None
    elif isinstance(hot, str):
        hot = store.hot_items(fs.m, fs.d, batch_size)
    a.item_slot, a.hot_items, a.n_hot = (ptr(hot[0]), ptr(hot[1]), hot[1].numel()) if hot else (None, None, 0)
    with torch.cuda.device(dev):
        a.stream = current_stream()
        check(lib.mfcd_train_epoch(C.byref(a)), "mfcd_train_epoch")
    fs.step += n_steps
    return losses[:n_steps]


def eval_batches(fs: _FlatState, store: TripletStore, batch_size):
    """-> (per-batch mean BCE as float32 device tensor, #correct as int)."""
    N = len(store)
    nb = (N + batch_size - 1) // batch_size
    dev = fs.params.device
    batch_loss = torch.zeros(max(nb, 1), dtype=torch.float32, device=dev)
    acc = torch.zeros(1 + max(nb, 1), dtype=torch.int64, device=dev)    # [#correct | fixed-point batch sums]
    correct = acc[:1]
    with torch.cuda.device(dev):
        check(lib.mfcd_triplet_eval(ptr(fs.U), ptr(fs.V), ptr(store.rec), N, fs.d, batch_size, ptr(batch_loss),
                                    ptr(correct), ptr(acc[1:]), current_stream()), "mfcd_triplet_eval")
    return batch_loss[:nb], correct


def _sum_like_python(values):
    """`total += loss.item()` over float32 values, in double, in order."""
    if isinstance(values, torch.Tensor) and values.is_cuda:
        wait_for_stream(values.device)
    tot = 0.0
    for v in (values if isinstance(values, list) else values.tolist()):
        tot += v
    return tot


def dist_world(world_size=None):
    """(rank, world) of the data-parallel job this process belongs to.  world_size None = what torchrun's
    environment says (WORLD_SIZE / RANK / LOCAL_RANK); > 1 initialises torch.distributed (NCCL) if needed."""
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else int(os.environ.get("WORLD_SIZE", "1"))
    world_size = int(world_size)
    if world_size <= 1:
        return 0, 1
    if not dist.is_initialized():
        from . import dist as mdist
        mdist.init_from_env()
    if not dist.is_initialized() or dist.get_world_size() != world_size:
        raise ValueError(f"world_size={world_size} needs a torch.distributed job of that size "
                         f"(launch with torchrun --nproc-per-node {world_size})")
    return dist.get_rank(), world_size


def _all_sum(values, dev, world):
    """sum of a few python numbers over the ranks (float64)"""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


class _Stager:
    """Two device staging slots + a copy stream for a HostTripletLoader (kept on the loader between epochs)."""

    def __init__(self, loader: HostTripletLoader, dev):
        self.dev = dev
        B = loader.batch_size
        self.rec = [torch.empty((B, 4), dtype=torch.int32, device=dev) for _ in range(2)] \
            if loader.fmt != "wire_rle" else [None, None]
        self.raw = None
        self.live = loader.fmt == "wire8_live"
        if loader.fmt in ("wire8", "wire8_live"):
            self.raw = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(2)]
        if self.live:
            # the packer's ring: slot k % 3 is packed (host threads) -> copied (DMA) -> free again
            from concurrent.futures import ThreadPoolExecutor
            self.hpk = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(3)]
            self.copied = [torch.cuda.Event() for _ in range(3)]
            self.packer = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mfcd-pack")
            self.fraction = self._balance(loader) if loader.pack_fraction == "auto" else float(loader.pack_fraction)
            loader._auto_fraction = self.fraction
        elif loader.fmt == "wire_rle":
            cap = max(b.numel() for b in loader.batches)
            self.raw = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]


def _stager_balance(self, loader):
    """pack_fraction='auto': time the packer on one whole batch (Tp) and the link on one raw batch (Tc), both
    alone; with a fraction f packed the packer needs f Tp and the link (1 - f / 2) Tc per step -> f = Tc / (Tp + Tc / 2)."""
    import time
    B = loader.batch_size
    if not loader.batches or B < 1:
        return 1.0
    src = loader.batches[0]
    n = int(loader.sizes[0])
    bad = C.c_int32(0)
    tp = tc = float("inf")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        t0 = time.perf_counter()
        check(lib.mfcd_host_pack_triplets8(src.data_ptr(), n, self.hpk[0].data_ptr(), loader.pack_threads,
                                           C.byref(bad)), "mfcd_host_pack_triplets8")
        tp = min(tp, time.perf_counter() - t0)
        a.record()
        self.rec[0][:n].copy_(src[:n], non_blocking=True)
        b.record()
        b.synchronize()
        tc = min(tc, a.elapsed_time(b) * 1e-3)
    f = tc / (tp + 0.5 * tc)
    return float(min(1.0, max(0.05, f)))


_Stager._balance = _stager_balance


def stager_for(loader: HostTripletLoader, dev) -> _Stager:
    st = loader.__dict__.get("_stager")
    if st is None or st.dev != torch.device(dev):
        with torch.cuda.device(dev):
            st = loader._stager = _Stager(loader, torch.device(dev))
    return st


def stream_epoch(fs: _FlatState, loader: HostTripletLoader, spec: OptimizerSpec, scatter, dp=None, hot=None):
    """One epoch over a HOST-resident loader: per optimiser step the batch is copied host -> device from pinned
    memory on a copy stream (double-buffered: the copy of batch k+1 overlaps the step on batch k), K1 runs on the
    staged batch (wire8: after the unpack kernel; wire_rle: K1 decodes the wire words itself), then the update /
    exchange, and the step's loss is copied back and READ BY THE HOST every step (one step behind the launches, so
    the GPU never waits for python).  Returns the per-step losses as a python list."""
    dev = fs.params.device
    st = stager_for(loader, dev)
    n_steps = len(loader)
    world = dp.world if dp is not None else 1
    sizes = torch.tensor(loader.sizes, dtype=torch.int64, device=dev)
    if world > 1:
        import torch.distributed as tdist
        tdist.all_reduce(sizes)                       # global batch of every step
    gsizes = sizes.tolist()
    losses = torch.zeros(max(n_steps, 1), dtype=torch.float32, device=dev)
    loss_host = torch.zeros(max(n_steps, 1), dtype=torch.float32).pin_memory()
    nU = fs.n * fs.d
    direct = loader.fmt == "wire_rle"
    flags = (_lib.FLAG_USER_GROUPED if loader.user_grouped else 0) | (_lib.FLAG_WIRE_RLE if direct else 0)
    slot, items, n_hot = (ptr(hot[0]), ptr(hot[1]), hot[1].numel()) if hot else (None, None, 0)
    out = []
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()

        packed = {}

        def pack(k):
            # worker thread: raw records of batch k -> 8-byte words in pinned slot k % 3, once the DMA that last
            # read that slot (batch k - 3) has finished; the C call releases the GIL and fans out over host threads
            s, h = k % 3, loader.batches[k]
            st.copied[s].synchronize()
            bad = C.c_int32(0)
            check(lib.mfcd_host_pack_triplets8(h.data_ptr(), n_packed(k), st.hpk[s].data_ptr(),
                                               loader.pack_threads, C.byref(bad)), "mfcd_host_pack_triplets8")
            if bad.value:
                raise _lib.MfcdError("wire8_live: soft labels or indices beyond 2^23 users / 2^20 items do not fit "
                                     "the 8-byte staging format (use fmt='records16')")

        def n_packed(k):
            # records of batch k that travel packed (the first ones; the rest goes raw in the same step)
            return loader.sizes[k] if st.fraction >= 1.0 else max(1, int(loader.sizes[k] * st.fraction))

        def submit_pack(k):
            if st.live and k < n_steps:
                packed[k] = st.packer.submit(pack, k)

        def upload(k):
            b = k % 2
            with torch.cuda.stream(st.copy_stream):
                st.copy_stream.wait_event(st.freed[b])
                h = loader.batches[k]
                if loader.fmt == "records16":
                    st.rec[b][: h.shape[0]].copy_(h, non_blocking=True)
                elif st.live:
                    n8, nk = n_packed(k), loader.sizes[k]
                    if n8 < nk:                      # the raw tail needs nothing from the packer: it goes first
                        st.rec[b][n8:nk].copy_(h[n8:nk], non_blocking=True)
                    packed.pop(k).result()
                    st.raw[b][:n8].copy_(st.hpk[k % 3][:n8], non_blocking=True)
                    st.copied[k % 3].record(st.copy_stream)
                    check(lib.mfcd_unpack_triplets8(ptr(st.raw[b]), n8, ptr(st.rec[b]),
                                                    st.copy_stream.cuda_stream), "mfcd_unpack_triplets8")
                else:
                    st.raw[b][: h.numel()].copy_(h, non_blocking=True)
                    if loader.fmt == "wire8":
                        check(lib.mfcd_unpack_triplets8(ptr(st.raw[b]), loader.sizes[k], ptr(st.rec[b]),
                                                        st.copy_stream.cuda_stream), "mfcd_unpack_triplets8")
                st.ready[b].record(st.copy_stream)
            submit_pack(k + 3)                       # its slot is the one this copy frees

        for b in range(2):
            st.freed[b].record(main)
        for k in range(3):
            submit_pack(k)
        if n_steps:
            upload(0)
        pending = None
        for k in range(n_steps):
            if k + 1 < n_steps and not st.live:
                upload(k + 1)
            b = k % 2
            main.wait_event(st.ready[b])
            src = st.raw[b] if direct else st.rec[b]
            Bk = loader.sizes[k]
            if scatter == MODE_ATOMIC:
                check(lib.mfcd_triplet_fwd_bwd_ex(ptr(fs.params), ptr(fs.params[nU:]), ptr(src), None, 0, Bk, fs.d,
                                                  1.0 / float(gsizes[k]), ptr(fs.grads), ptr(fs.grads[nU:]),
                                                  ptr(losses[k:k + 1]), slot, items, n_hot, flags, main.cuda_stream),
                      "mfcd_triplet_fwd_bwd_ex")
            else:
                if direct:
                    raise _lib.MfcdError("the wire_rle staging format is decoded by the atomic K1 only")
                need = C.c_size_t(0)
                check(lib.mfcd_det_workspace_bytes_nm(Bk, fs.d, fs.n, fs.m, C.byref(need)), "mfcd_det_workspace_bytes_nm")
                ws = fs.ensure_workspace(need.value)
                check(lib.mfcd_triplet_fwd_bwd_det(ptr(fs.params), ptr(fs.params[nU:]), ptr(src), None, 0, Bk, fs.d,
                                                   1.0 / float(gsizes[k]), fs.n, fs.m, ptr(fs.grads),
                                                   ptr(fs.grads[nU:]), ptr(losses[k:k + 1]), ptr(ws),
                                                   ws.numel() if ws is not None else 0, main.cuda_stream),
                      "mfcd_triplet_fwd_bwd_det")
            st.freed[b].record(main)
            if k + 1 < n_steps and st.live:
                upload(k + 1)                        # waits for the packer: after this step's K1 is queued
            step = fs.step + 1
            if dp is not None and dp.exchange is not None:
                dp.exchange.step(fs, spec, step)
            else:
                if dp is not None:
                    import torch.distributed as tdist
                    tdist.all_reduce(fs.grads)
                if spec.kind == 0:
                    check(lib.mfcd_adam_update(ptr(fs.params), ptr(fs.grads), ptr(fs.state1), ptr(fs.state2),
                                               fs.params.numel(), spec.lr, spec.beta1, spec.beta2, spec.eps,
                                               spec.weight_decay, step, 1, main.cuda_stream), "mfcd_adam_update")
                else:
                    check(lib.mfcd_sgd_update(ptr(fs.params), ptr(fs.grads), ptr(fs.state1), fs.params.numel(), spec.lr,
                                              spec.momentum, spec.weight_decay, step, 1, main.cuda_stream),
                          "mfcd_sgd_update")
            fs.step = step
            # the step's result goes back to the host every step; it is read once the NEXT step is queued
            loss_host[k:k + 1].copy_(losses[k:k + 1], non_blocking=True)
            st.done[b].record(main)
            if pending is not None:
                st.done[pending % 2].synchronize()
                out.append(float(loss_host[pending]))
            pending = k
        if pending is not None:
            st.done[pending % 2].synchronize()
            out.append(float(loss_host[pending]))
    if world > 1:       # local partial sums (each scaled by 1 / global batch) -> global batch means
        t = torch.tensor(out, dtype=torch.float64, device=dev)
        import torch.distributed as tdist
        tdist.all_reduce(t)
        out = t.tolist()
        if dp.exchange is not None and dp.exchange.sync == "kernel":
            dp.exchange.check_error()
    return out


def train_epoch(fs: _FlatState, train_loader: TripletLoader, spec: OptimizerSpec, scatter, dp=None):
    """One training epoch through the product path: the epoch's reshuffle, then every optimiser step
    (K1 -> [gradient exchange] -> K3), no host sync.  Returns the per-step batch-mean losses (float32, device).

    Throughput batches (atomic scatter, more than 256 triplets): the reshuffle is the one-pass multisplit of
    csrc/epoch_batches.cu -- same batches as walking a random permutation in chunks of B, each batch laid out
    grouped by user -- so K1 runs with user runs and hot-row privatisation.  Reference-sized or deterministic
    batches: the epoch permutation is handed to the kernels as is (batch order = the reference's summation order).
    dp: a dist.DataParallel (replicated tables, this rank's shard in `train_loader`)."""
    if isinstance(train_loader, HostTripletLoader):
        return stream_epoch(fs, train_loader, spec, scatter, dp=dp, hot=train_loader.__dict__.get("hot"))
    if dp is not None:
        return dp.train_epoch(train_loader, scatter)
    B = train_loader.batch_size
    if scatter == MODE_ATOMIC and B > 256:
        er = train_loader.epoch_records(B)
        if er is not None:
            ep_store, flags = er
            hot = train_loader.store.hot_items(fs.m, fs.d, B)
            return run_epoch(fs, ep_store, None, B, spec, scatter, flags=flags, hot=hot)
    perm = train_loader.epoch_perm()
    return run_epoch(fs, train_loader.store, perm, B, spec, scatter)


def train_model(model, train_loader, val_loader, optimizer, device, num_epochs=100, is_last=False,
                open_browser=False, *, mode="auto", progress=False, world_size=None, dp_backend="peer"):
    """Same contract as the reference's train_model (structure.py:812-878):
    returns ``(train_losses, val_losses)``, one mean-of-batch-means per epoch.

    The whole training runs without a host sync (the reference syncs on ``loss.item()`` every step); the per-epoch
    losses are read once at the end.
    Keyword-only extras (reference behaviour by default): ``mode`` = scatter mode ('auto': deterministic up to
    256 triplets per batch, atomic above); ``world_size`` > 1 (or a torchrun environment) = data-parallel training,
    every rank passing ITS shard of the triplets and ``train_loader.batch_size`` meaning the GLOBAL batch."""
    dev = compute_device(device)
    train_loader = as_loader(train_loader, dev)
    val_loader = as_loader(val_loader, dev)
    spec = OptimizerSpec(optimizer)
    rank, world = dist_world(world_size)
    dp = None
    if world > 1:
        from . import dist as mdist
        dp = mdist.DataParallel.attach(model, dev, spec, backend=dp_backend)
        fs = dp.fs
    else:
        fs = model.flat_state(dev)
    _import_optimizer_state(model, optimizer, fs)
    scatter = resolve_mode(mode, train_loader.batch_size)

    train_losses, val_losses = [], []
    epochs = range(num_epochs)
    if progress and rank == 0:
        from tqdm import tqdm
        epochs = tqdm(epochs, desc="Training Progress")
    pending = []                                       # single process: per-epoch device tensors, read after the loop
    for _ in epochs:
        step_losses = train_epoch(fs, train_loader, spec, scatter, dp=dp)
        val_loader.begin_iteration()
        vloss, _ = eval_batches(fs, val_loader.store, val_loader.batch_size)
        if world > 1:       # step_losses already holds the GLOBAL batch means (all-reduced once per epoch)
            train_losses.append(_sum_like_python(step_losses) / len(step_losses))
            vsum, vcnt = _all_sum([_sum_like_python(vloss), len(val_loader)], dev, world)
            val_losses.append(vsum / vcnt)
        else:
            pending.append((step_losses, vloss))
    if pending:
        # ONE host sync for the whole training (the reference syncs on loss.item() every step): the epochs run back to
        # back on the GPU, and the wait happens in Stream.synchronize, which releases the GIL -- a sweep worker that
        # waited inside Tensor.tolist() kept the caller thread of a concurrent sweep from preparing the next repetition
        wait_for_stream(dev)
        for step_losses, vloss in pending:
            train_losses.append(_sum_like_python(step_losses) / len(train_loader))
            val_losses.append(_sum_like_python(vloss) / len(val_loader))      # ZeroDivisionError like the reference
    _export_optimizer_state(model, optimizer, fs, spec)
    if device is not None and torch.device(device).type == "cpu":
        model.mirror_to_host()
    return train_losses, val_losses


def evaluate_model(model, test_loader, device, *, world_size=None):
    """(mean-of-batch-means BCE, accuracy) on the test loader (structure.py:881-921).  Data parallel: every
    rank passes its shard; sums are combined over the ranks."""
    dev = compute_device(device)
    loader = as_loader(test_loader, dev)
    fs = model.flat_state(dev)
    rank, world = dist_world(world_size)
    loader.begin_iteration()
    batch_loss, correct = eval_batches(fs, loader.store, loader.batch_size)
    lsum, nb, ncorrect, total = _all_sum([_sum_like_python(batch_loss), len(loader), int(correct.item()),
                                          len(loader.store)], dev, world)
    accuracy = ncorrect / total if total > 0 else 0.0
    return lsum / nb, accuracy


def compute_ground_truth_metrics(test_loader, X, device, *, world_size=None):
    """(mean-of-batch-means MSE of sigmoid(X[u,i]-X[u,j]) vs label, accuracy of
    (diff > 0) == label) -- structure.py:1085-1127; note: no scale s."""
    dev = compute_device(device)
    rank, world = dist_world(world_size)
    loader = as_loader(test_loader, dev)
    gt = GroundTruth.wrap(X, dev)
    loader.begin_iteration()
    N = len(loader.store)
    nb = len(loader)
    batch_mse = torch.zeros(max(nb, 1), dtype=torch.float32, device=dev)
    acc = torch.zeros(1 + max(nb, 1), dtype=torch.int64, device=dev)    # [#correct | fixed-point batch sums]
    correct = acc[:1]
    xv = gt.xview()
    with torch.cuda.device(dev):
        check(lib.mfcd_ground_truth_eval(C.byref(xv), ptr(loader.store.rec), N, loader.batch_size, ptr(batch_mse),
                                         ptr(correct), ptr(acc[1:]), current_stream()), "mfcd_ground_truth_eval")
    msum, nbs, ncorrect, total = _all_sum([_sum_like_python(batch_mse[:nb]), nb, int(correct.item()), N], dev, world)
    accuracy = ncorrect / total if total > 0 else 0.0
    return msum / nbs, accuracy
