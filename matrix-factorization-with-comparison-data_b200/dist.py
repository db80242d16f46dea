"""Data-parallel training: one process per GPU, tables replicated, triplets
sharded, dense gradients combined with a bucketed all-reduce each step.

The reference is single-process; this is the new
multi-GPU layer of section 8(e).  Per optimiser step every rank runs K1 on its
slice of the global batch with ``inv_batch = 1 / B_global`` (so the summed
gradient is the global batch-mean gradient), the flat gradient buffer
``(n+m)*d`` fp32 is all-reduced (NCCL over NVLink/NVSwitch) in buckets, and each
bucket is consumed by the fused Adam kernel as soon as it lands, overlapping the
next bucket's transfer.  Every rank applies the identical update, so replicas
stay bit-identical.  Per-step losses are all-reduced once per epoch.

``engine`` abstracts the two device operations so the host logic can be
exercised with world_size-2 gloo on CPU (tests inject an oracle engine); the
product engine is ``CudaEngine`` (C ABI calls only, no fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import lib, check, ptr, current_stream


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def split_even(total, parts, index):
    """[a, b) of part ``index`` when ``total`` items are split into ``parts`` nearly equal runs."""
    base, rem = divmod(total, parts)
    a = index * base + min(index, rem)
    return a, a + base + (1 if index < rem else 0)


class ReplicatedPlan:
    """Every rank holds the whole dataset and the same epoch order; global batch k
    is positions [k*B, k*B + B_k) and rank r takes its even slice of it.  Exactly
    the batches a single GPU with batch size B would see."""

    def __init__(self, n_samples, global_batch, rank, world):
        self.N, self.B, self.rank, self.world = n_samples, global_batch, rank, world

    def n_steps(self):
        return (self.N + self.B - 1) // self.B

    def local_range(self, k):
        start = k * self.B
        bk = min(self.B, self.N - start)
        a, b = split_even(bk, self.world, self.rank)
        return start + a, b - a, bk


class PartitionedPlan:
    """Each rank holds only its shard (sizes known to all); global batch k is the
    union of every rank's local positions [k*Bl, (k+1)*Bl)."""

    def __init__(self, shard_sizes, local_batch, rank):
        self.sizes, self.Bl, self.rank = list(shard_sizes), local_batch, rank

    def n_steps(self):
        return max((s + self.Bl - 1) // self.Bl for s in self.sizes)

    def _local(self, r, k):
        start = k * self.Bl
        return start, max(0, min(self.Bl, self.sizes[r] - start))

    def local_range(self, k):
        start, bl = self._local(self.rank, k)
        total = sum(self._local(r, k)[1] for r in range(len(self.sizes)))
        return start, bl, total


class CudaEngine:
    """K1 / K3 on this rank's GPU through the C ABI."""

    def __init__(self, flat_state, store, perm, spec, mode, use_hot=True, flags=None, hot="auto"):
        self.fs, self.store, self.perm, self.spec, self.mode = flat_state, store, perm, spec, mode
        self.use_hot = use_hot
        self.flags, self.hot = flags, hot       # overrides: K1 flags of the batch layout, (item_slot, hot_items)
        self.ws = None

    def grads(self):
        return self.fs.grads

    def fwd_bwd(self, start, b_local, b_global, loss_slot):
        fs = self.fs
        if b_local == 0:
            return
        nU = fs.n * fs.d
        inv = 1.0 / float(b_global)
        if self.mode == 1:
            need = C.c_size_t(0)
            check(lib.mfcd_det_workspace_bytes_nm(b_local, fs.d, fs.n, fs.m, C.byref(need)), "mfcd_det_workspace_bytes_nm")
            ws = fs.ensure_workspace(need.value)
            check(lib.mfcd_triplet_fwd_bwd_det(ptr(fs.params), ptr(fs.params[nU:]), ptr(self.store.rec),
                                               ptr(self.perm), start, b_local, fs.d, inv, fs.n, fs.m,
                                               ptr(fs.grads), ptr(fs.grads[nU:]), ptr(loss_slot), ptr(ws),
                                               ws.numel() if ws is not None else 0, current_stream()),
                  "mfcd_triplet_fwd_bwd_det")
        else:
            hot = self.hot
            if not self.use_hot:
                hot = None
            elif isinstance(hot, str):
                hot = self.store.hot_items(fs.m, fs.d, b_local)
            slot, items, n_hot = (ptr(hot[0]), ptr(hot[1]), hot[1].numel()) if hot else (None, None, 0)
            flags = self.store.k1_flags(b_local, self.perm) if self.flags is None else int(self.flags)
            check(lib.mfcd_triplet_fwd_bwd_ex(ptr(fs.params), ptr(fs.params[nU:]), ptr(self.store.rec),
                                              ptr(self.perm), start, b_local, fs.d, inv, ptr(fs.grads),
                                              ptr(fs.grads[nU:]), ptr(loss_slot), slot, items, n_hot,
                                              flags, current_stream()),
                  "mfcd_triplet_fwd_bwd_ex")

    def update(self, a, b, step):
        fs, s = self.fs, self.spec
        if s.kind == 0:
            check(lib.mfcd_adam_update(ptr(fs.params[a:]), ptr(fs.grads[a:]), ptr(fs.state1[a:]), ptr(fs.state2[a:]),
                                       b - a, s.lr, s.beta1, s.beta2, s.eps, s.weight_decay, step, 1,
                                       current_stream()), "mfcd_adam_update")
        else:
            check(lib.mfcd_sgd_update(ptr(fs.params[a:]), ptr(fs.grads[a:]), ptr(fs.state1[a:]), b - a, s.lr,
                                      s.momentum, s.weight_decay, step, 1, current_stream()), "mfcd_sgd_update")


class PeerExchange:
    """Symmetric (peer-mapped) parameter and gradient buffers + the fused exchange kernel K9.

    torch.distributed's symmetric memory is used for what it is -- allocation, handle exchange and
    cross-rank barriers (plumbing); the data movement and the arithmetic are mfcd_dp_fused_adam.
    Use:  ex = PeerExchange(numel, dev);  fs = model.flat_state(dev, storage=ex.storage());
          per step: K1 into fs.grads, then ex.step(fs, spec, step)."""

    def __init__(self, numel, device, group=None, use_multimem="auto", sync=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        padded = (numel + 3) // 4 * 4
        self.numel = numel
        self.params = symm_mem.empty(padded, dtype=torch.float32, device=device)
        self.grads = symm_mem.empty(padded, dtype=torch.float32, device=device)
        # second gradient buffer (in-kernel sync): step k accumulates into buffer k & 1 while K9 clears the other one
        # LOCALLY -- no zeros through the fabric (they were as much NVLink traffic again as the all-gather)
        self.grads_b = symm_mem.empty(padded, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)      # ready[8] | done[8] (+ padding)
        self.params.zero_()
        self.grads.zero_()
        self.grads_b.zero_()
        self.flags.zero_()
        self.hp = symm_mem.rendezvous(self.params, group)
        self.hg = symm_mem.rendezvous(self.grads, group)
        self.hg_b = symm_mem.rendezvous(self.grads_b, group)
        self.hf = symm_mem.rendezvous(self.flags, group)
        self.rank, self.world = self.hp.rank, self.hp.world_size
        self.peer_params = (C.c_uint64 * self.world)(*[int(x) for x in self.hp.buffer_ptrs])
        self.peer_grads = (C.c_uint64 * self.world)(*[int(x) for x in self.hg.buffer_ptrs])
        self.peer_grads_b = (C.c_uint64 * self.world)(*[int(x) for x in self.hg_b.buffer_ptrs])
        self.peer_flags = (C.c_uint64 * self.world)(*[int(x) for x in self.hf.buffer_ptrs])
        mc_p = int(getattr(self.hp, "multicast_ptr", 0) or 0)
        mc_g = int(getattr(self.hg, "multicast_ptr", 0) or 0)
        mc_gb = int(getattr(self.hg_b, "multicast_ptr", 0) or 0)
        # measured on NVSwitch B200 boxes (profiles/r01_notes.md): in-switch multimem reduction wins at 8 ranks
        # (0.134 vs 0.163 ms per exchange), plain peer loads win at 2 (0.094 vs 0.139 ms); equal at 4.
        env = os.environ.get("MFCD_DP_MULTIMEM", "auto")
        if env in ("on", "off"):
            use_multimem = env == "on"
        if use_multimem == "auto":
            use_multimem = self.world > 4
        self.multimem = bool(use_multimem and mc_p and mc_g and mc_gb)
        self.mc_params, self.mc_grads, self.mc_grads_b = (mc_p, mc_g, mc_gb) if self.multimem else (0, 0, 0)
        self.double_buffer = os.environ.get("MFCD_DP_DOUBLE_BUFFER", "1") != "0"
        # "kernel": flags exchanged inside K9, the owner clears the gradient slices it consumed (default);
        # "barrier": two symmetric-memory barriers around K9 + a local memset (round-1 path, kept for comparison)
        self.sync = sync or os.environ.get("MFCD_DP_SYNC", "kernel")
        self.seq = 0
        self.cta_counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.hp.barrier(channel=0)

    def storage(self):
        return self.params, self.grads

    def shard(self):
        b, e = C.c_int64(0), C.c_int64(0)
        check(lib.mfcd_dp_shard_range(self.numel, self.rank, self.world, C.byref(b), C.byref(e)), "mfcd_dp_shard_range")
        return b.value, e.value

    def check_error(self):
        """one host sync: raises if a bounded in-kernel wait timed out since the last check"""
        if int(self.error.item()) != 0:
            self.error.zero_()
            raise _lib.MfcdError("mfcd_dp_fused_adam_sync: a peer did not arrive in time (results are invalid)")

    def step(self, fs, spec, step):
        """K9: reduce-scatter + Adam + all-gather in one kernel; leaves the gradient buffers cleared."""
        if spec.kind != 0:
            raise NotImplementedError("the fused peer exchange implements Adam")
        assert fs.params.data_ptr() == self.params.data_ptr()
        if self.sync == "kernel":
            self.seq += 1
            in_a = fs.grads.data_ptr() == self.grads.data_ptr()
            assert in_a or fs.grads.data_ptr() == self.grads_b.data_ptr()
            peers, mc = (self.peer_grads, self.mc_grads) if in_a else (self.peer_grads_b, self.mc_grads_b)
            other = (self.grads_b if in_a else self.grads) if self.double_buffer else None
            check(lib.mfcd_dp_fused_adam_sync(peers, self.peer_params, self.peer_flags, mc,
                                              self.mc_params, self.rank, self.world, self.numel, ptr(fs.state1),
                                              ptr(fs.state2), spec.lr, spec.beta1, spec.beta2, spec.eps,
                                              spec.weight_decay, step, self.seq, ptr(self.cta_counter),
                                              ptr(self.error), ptr(other), current_stream()), "mfcd_dp_fused_adam_sync")
            if other is not None:           # the next K1 accumulates into the buffer this call has just cleared
                fs.grads = other[:fs.grads.numel()]
            return
        assert fs.grads.data_ptr() == self.grads.data_ptr()
        self.hg.barrier(channel=0)
        check(lib.mfcd_dp_fused_adam(self.peer_grads, self.peer_params, self.mc_grads, self.mc_params, self.rank,
                                     self.world, self.numel, ptr(fs.state1), ptr(fs.state2), spec.lr, spec.beta1,
                                     spec.beta2, spec.eps, spec.weight_decay, step, current_stream()),
              "mfcd_dp_fused_adam")
        self.hp.barrier(channel=1)
        fs.grads.zero_()


def dp_step_peer(engine, plan, k, step, loss_slot, exchange):
    """Data-parallel step with the fused peer-memory exchange instead of NCCL all-reduce + K3."""
    start, b_local, b_global = plan.local_range(k)
    engine.fwd_bwd(start, b_local, b_global, loss_slot)
    exchange.step(engine.fs, engine.spec, step)


def bucket_bounds(numel, bucket_elems):
    """16-byte aligned bucket edges covering [0, numel)."""
    bucket_elems = max(4, (bucket_elems // 4) * 4)
    edges = list(range(0, numel, bucket_elems)) + [numel]
    return list(zip(edges[:-1], edges[1:]))


def dp_step(engine, plan, k, step, loss_slot, group=None, bucket_elems=0):
    """One data-parallel optimiser step: local K1 -> all-reduce of the flat gradient -> K3.
    bucket_elems == 0 (default): ONE all-reduce (38 MB at config 4: 0.09 ms on 2 NVLinked B200s, measured
    faster than 16 MB buckets, whose per-call latency outweighs the overlap with K3); > 0: buckets, each
    consumed by K3 as soon as it lands."""
    start, b_local, b_global = plan.local_range(k)
    engine.fwd_bwd(start, b_local, b_global, loss_slot)
    g = engine.grads()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1 and bucket_elems <= 0:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        engine.update(0, g.numel(), step)
        return
    bounds = bucket_bounds(g.numel(), bucket_elems if bucket_elems > 0 else g.numel())
    if world > 1:
        works = [dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True) for a, b in bounds]
        for (a, b), w in zip(bounds, works):
            w.wait()                       # current stream waits for this bucket only
            engine.update(a, b, step)
    else:
        engine.update(0, g.numel(), step)


def dp_epoch(engine, plan, step0, losses, group=None, bucket_elems=0, exchange=None):
    """All steps of one epoch; ``losses`` (n_steps floats on the engine's device,
    zero-filled) receives the global batch-mean loss of every step.  ``exchange``: a PeerExchange
    to use the fused peer-memory path instead of NCCL all-reduce + K3."""
    n_steps = plan.n_steps()
    for k in range(n_steps):
        if exchange is not None:
            dp_step_peer(engine, plan, k, step0 + k + 1, losses[k:k + 1], exchange)
        else:
            dp_step(engine, plan, k, step0 + k + 1, losses[k:k + 1], group=group, bucket_elems=bucket_elems)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group)
    if exchange is not None and exchange.sync == "kernel":
        exchange.check_error()
    return n_steps


class DataParallel:
    """What train_model drives under world_size > 1: replicated tables (in peer-mapped symmetric memory for the
    fused K9 exchange), this rank's shard of the triplets, the global batch split evenly over the ranks."""

    def __init__(self, model, device, spec, backend="peer"):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device, self.backend, self.spec = device, backend, spec
        n, d = model.U.shape
        m = model.V.shape[0]
        self.exchange = None
        if backend == "peer" and spec.kind == 0:
            self.exchange = PeerExchange((n + m) * d, device)
            self.fs = model.flat_state(device, storage=self.exchange.storage())
        else:
            self.fs = model.flat_state(device)
        # replicas start bit-identical whatever each rank's host RNG has been through
        dist.broadcast(self.fs.params, src=0)

    @classmethod
    def attach(cls, model, device, spec, backend="peer"):
        dp = getattr(model, "_dp", None)
        if dp is None or dp.device != device or dp.backend != backend or model._flat is not dp.fs \
                or (dp.exchange is None) != (backend != "peer" or spec.kind != 0):
            dp = cls(model, device, spec, backend)
            model._dp = dp
        dp.spec = spec
        return dp

    def train_epoch(self, loader, scatter):
        """-> per-step GLOBAL batch-mean losses (float32 device tensor), identical on every rank"""
        fs = self.fs
        B = loader.batch_size
        Bl = (B + self.world - 1) // self.world
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        sizes[self.rank] = len(loader.store)
        dist.all_reduce(sizes)
        plan = PartitionedPlan(sizes.tolist(), Bl, self.rank)
        n_steps = plan.n_steps()
        losses = torch.zeros(max(n_steps, 1), dtype=torch.float32, device=self.device)
        store, perm, flags, hot = loader.store, None, None, "auto"
        er = loader.epoch_records(Bl) if (scatter == 0 and Bl > 256) else None
        if er is not None:
            store, flags = er
            hot = loader.store.hot_items(fs.m, fs.d, Bl)
        else:
            perm = loader.epoch_perm()
        engine = CudaEngine(fs, store, perm, self.spec, scatter, flags=flags, hot=hot)
        with torch.cuda.device(self.device):
            dp_epoch(engine, plan, fs.step, losses, exchange=self.exchange)
        fs.step += n_steps
        return losses[:n_steps]
