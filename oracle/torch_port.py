"""Torch-CPU port of the reference's training / evaluation step.

TEST INFRASTRUCTURE ONLY (same rules as mfcd_oracle.py): used by tests, by
``__graft_entry__.smoke()`` and by the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the CPU implementation that is timed on the GPU box's
host cores.  The reference is pure Python on top of PyTorch CPU ops and cannot
travel to the GPU box (/root/reference does not exist there), so this module
restates its hot loop with the SAME ATen operators in the same order -- it
executes the same kernels the reference executes -- citing the lines it follows.
Pinned against the reference by tests/test_oracle_vs_golden.py::test_torch_port_*.
"""
from __future__ import annotations

import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class PortModel(nn.Module):
    """structure.py:766-795."""

    def __init__(self, U, V):
        super().__init__()
        self.U = nn.Parameter(U.clone())
        self.V = nn.Parameter(V.clone())

    def forward(self, u, i, j):
        diff = torch.sum(self.U[u] * (self.V[i] - self.V[j]), dim=1)
        return torch.sigmoid(diff)


def make_optimizer(model, lr, weight_decay):
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)     # structure.py:364


def train_step(model, optimizer, u, i, j, z):
    """structure.py:847-852 for one batch; returns loss.item()."""
    optimizer.zero_grad()
    pred = model(u, i, j)
    loss = F.binary_cross_entropy(pred, z.float())
    loss.backward()
    optimizer.step()
    return loss.item()


def eval_batches(model, batches):
    """structure.py:896-921 -> (mean of batch-mean BCE, accuracy)."""
    tot, correct, count = 0.0, 0, 0
    with torch.no_grad():
        for (u, i, j, z) in batches:
            pred = model(u, i, j)
            tot += F.binary_cross_entropy(pred, z.float()).item()
            correct += ((pred > 0.5).float() == z).sum().item()
            count += len(z)
    return tot / len(batches), (correct / count if count else 0.0)


def time_train_steps(n, m, d, batch, steps, warmup=2, lr=1e-3, weight_decay=1e-5, seed=0, threads=None,
                     item_probs=None):
    """Time `steps` reference-equivalent optimiser steps on synthetic triplets of the given shape.
    Returns (triplets_per_second, seconds_per_step, threads_used)."""
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    scale = torch.sqrt(torch.tensor(d, dtype=torch.float32))
    model = PortModel(torch.randn(n, d, generator=g) / scale, torch.randn(m, d, generator=g) / scale)
    opt = make_optimizer(model, lr, weight_decay)
    total = steps + warmup
    u = torch.randint(0, n, (total, batch), generator=g)
    if item_probs is None:
        i = torch.randint(0, m, (total, batch), generator=g)
        j = torch.randint(0, m, (total, batch), generator=g)
    else:
        i = torch.multinomial(item_probs, total * batch, replacement=True, generator=g).view(total, batch)
        j = torch.multinomial(item_probs, total * batch, replacement=True, generator=g).view(total, batch)
    z = torch.randint(0, 2, (total, batch), generator=g).double()
    for k in range(warmup):
        train_step(model, opt, u[k], i[k], j[k], z[k])
    t0 = time.perf_counter()
    for k in range(warmup, total):
        train_step(model, opt, u[k], i[k], j[k], z[k])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()
