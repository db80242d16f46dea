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

    def __init__(self, flat_state, store, perm, spec, mode, use_hot=True):
        self.fs, self.store, self.perm, self.spec, self.mode = flat_state, store, perm, spec, mode
        self.use_hot = use_hot
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
            check(lib.mfcd_det_workspace_bytes(b_local, fs.d, C.byref(need)), "mfcd_det_workspace_bytes")
            ws = fs.ensure_workspace(need.value)
            check(lib.mfcd_triplet_fwd_bwd_det(ptr(fs.params), ptr(fs.params[nU:]), ptr(self.store.rec),
                                               ptr(self.perm), start, b_local, fs.d, inv, fs.n, fs.m,
                                               ptr(fs.grads), ptr(fs.grads[nU:]), ptr(loss_slot), ptr(ws),
                                               ws.numel() if ws is not None else 0, current_stream()),
                  "mfcd_triplet_fwd_bwd_det")
        else:
            hot = self.store.hot_items(fs.m, fs.d, b_local) if self.use_hot else None
            slot, items, n_hot = (ptr(hot[0]), ptr(hot[1]), hot[1].numel()) if hot else (None, None, 0)
            check(lib.mfcd_triplet_fwd_bwd_hot(ptr(fs.params), ptr(fs.params[nU:]), ptr(self.store.rec),
                                               ptr(self.perm), start, b_local, fs.d, inv, ptr(fs.grads),
                                               ptr(fs.grads[nU:]), ptr(loss_slot), slot, items, n_hot,
                                               current_stream()), "mfcd_triplet_fwd_bwd_hot")

    def update(self, a, b, step):
        fs, s = self.fs, self.spec
        if s.kind == 0:
            check(lib.mfcd_adam_update(ptr(fs.params[a:]), ptr(fs.grads[a:]), ptr(fs.state1[a:]), ptr(fs.state2[a:]),
                                       b - a, s.lr, s.beta1, s.beta2, s.eps, s.weight_decay, step, 1,
                                       current_stream()), "mfcd_adam_update")
        else:
            check(lib.mfcd_sgd_update(ptr(fs.params[a:]), ptr(fs.grads[a:]), ptr(fs.state1[a:]), b - a, s.lr,
                                      s.momentum, s.weight_decay, step, 1, current_stream()), "mfcd_sgd_update")


def bucket_bounds(numel, bucket_elems):
    """16-byte aligned bucket edges covering [0, numel)."""
    bucket_elems = max(4, (bucket_elems // 4) * 4)
    edges = list(range(0, numel, bucket_elems)) + [numel]
    return list(zip(edges[:-1], edges[1:]))


def dp_step(engine, plan, k, step, loss_slot, group=None, bucket_elems=4 << 20):
    """One data-parallel optimiser step: local K1 -> bucketed all-reduce -> K3 per bucket."""
    start, b_local, b_global = plan.local_range(k)
    engine.fwd_bwd(start, b_local, b_global, loss_slot)
    g = engine.grads()
    bounds = bucket_bounds(g.numel(), bucket_elems)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        works = [dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True) for a, b in bounds]
        for (a, b), w in zip(bounds, works):
            w.wait()                       # current stream waits for this bucket only
            engine.update(a, b, step)
    else:
        engine.update(0, g.numel(), step)


def dp_epoch(engine, plan, step0, losses, group=None, bucket_elems=4 << 20):
    """All steps of one epoch; ``losses`` (n_steps floats on the engine's device,
    zero-filled) receives the global batch-mean loss of every step."""
    n_steps = plan.n_steps()
    for k in range(n_steps):
        dp_step(engine, plan, k, step0 + k + 1, losses[k:k + 1], group=group, bucket_elems=bucket_elems)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group)
    return n_steps
