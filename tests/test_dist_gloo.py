"""world_size-2 gloo run of the data-parallel host logic on CPU.

The engine here is the numpy oracle (tests may call it); the product engine is
CudaEngine.  Checks: 2-rank DP with global batch B == 1-process training with
batch B, replicas stay identical, per-step losses are the global batch means."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, batches_from


class OracleEngine:
    def __init__(self, U, V, u, i, j, z, lr, wd):
        from oracle import mfcd_oracle as O
        self.O = O
        self.n, self.m, self.d = U.shape[0], V.shape[0], U.shape[1]
        self.params = torch.from_numpy(np.concatenate([U.ravel(), V.ravel()]).astype(np.float32))
        self._grads = torch.zeros_like(self.params)
        self.m1 = np.zeros(self.params.numel(), np.float32)
        self.m2 = np.zeros(self.params.numel(), np.float32)
        self.u, self.i, self.j, self.z = u, i, j, z.astype(np.float32)
        self.lr, self.wd = lr, wd

    def grads(self):
        return self._grads

    def tables(self):
        p = self.params.numpy()
        return p[: self.n * self.d].reshape(self.n, self.d), p[self.n * self.d:].reshape(self.m, self.d)

    def fwd_bwd(self, start, b_local, b_global, loss_slot):
        if b_local == 0:
            return
        O = self.O
        U, V = self.tables()
        sl = slice(start, start + b_local)
        p = O.forward(U, V, self.u[sl], self.i[sl], self.j[sl])
        loss_slot += float(O.bce_per_sample(p, self.z[sl]).sum()) / b_global
        gx = O.bce_grad_score(p, self.z[sl], b_global)
        gU, gV = O.dense_grads(U, V, self.u[sl], self.i[sl], self.j[sl], gx)
        self._grads += torch.from_numpy(np.concatenate([gU.ravel(), gV.ravel()]))

    def update(self, a, b, step):
        p = self.params.numpy()
        g = self._grads.numpy()
        self.O.adam_step(p[a:b], g[a:b].copy(), self.m1[a:b], self.m2[a:b], step, lr=self.lr, weight_decay=self.wd)
        g[a:b] = 0


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mfcd_b200 import dist as mdist
    g = load_golden("train_d10_k3.npz")
    u, i, j, z = g["batch_u"], g["batch_i"], g["batch_j"], g["batch_z"]
    spe = int(g["steps_per_epoch"])
    N = int(g["batch_sizes"][:spe].sum())                     # first epoch, already in the reference's order
    eng = OracleEngine(g["U0"].copy(), g["V0"].copy(), u[:N], i[:N], j[:N], z[:N], float(g["lr"]), float(g["wd"]))
    plan = mdist.ReplicatedPlan(N, 64, rank, world)
    losses = torch.zeros(plan.n_steps(), dtype=torch.float32)
    mdist.dp_epoch(eng, plan, 0, losses, bucket_elems=256)
    U, V = eng.tables()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), U=U, V=V, losses=losses.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_dp_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["U"], r1["U"]) and np.array_equal(r0["V"], r1["V"])      # replicas identical
    assert np.array_equal(r0["losses"], r1["losses"])
    # against the reference's own single-process run of the same first epoch
    g = load_golden("train_d10_k3.npz")
    spe = int(g["steps_per_epoch"])
    ref_losses = g["step_losses"][:spe]
    assert np.abs(r0["losses"] - ref_losses).max() < 1e-5 * np.abs(ref_losses).max()
    from oracle import mfcd_oracle as O
    U, V = g["U0"].copy(), g["V0"].copy()
    O.train_steps(U, V, batches_from(g)[:spe], float(g["lr"]), float(g["wd"]))
    assert np.abs(r0["U"] - U).max() < 1e-5 * np.abs(U).max()
    assert np.abs(r0["V"] - V).max() < 1e-5 * np.abs(V).max()
