"""Data-parallel training THROUGH THE PRODUCT API (trainer.train_model(world_size=2)), checked against the numpy
oracle; run under torchrun with 2 ranks:

  * MFCD_DP_TEST=gloo1  both ranks share cuda:0, torch.distributed over gloo, gradient exchange = all-reduce of
                        the CUDA-computed flat gradient + K3 (dp_backend="allreduce"): exercises DataParallel,
                        dp_epoch, the plans and the sharded evaluation on a ONE-GPU box;
  * MFCD_DP_TEST=nccl   one rank per GPU, NCCL, the fused peer-memory exchange K9 with in-kernel flags.

1. deterministic scatter, loaders walked in order: per-step losses, final U / V and validation loss equal the oracle
   run on the same GLOBAL batches (batch k = every rank's k-th local slice) to 1e-5;
2. atomic scatter at a throughput batch size with the per-epoch device reshuffle + user grouping: the oracle replays
   the exact batches (each rank's epoch seed and user-sorted shard are gathered) to 2e-5;
3. evaluate_model(world_size=2) over sharded test data == oracle over the union; replicas bit-identical.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KIND = os.environ.get("MFCD_DP_TEST", "nccl")
if KIND == "gloo1":
    os.environ["LOCAL_RANK"] = "0"          # both ranks compute on cuda:0

import mfcd_b200  # noqa: E402
from mfcd_b200 import trainer  # noqa: E402
from mfcd_b200.store import TripletLoader, TripletStore  # noqa: E402
from mfcd_b200.trainer import MatrixFactorization  # noqa: E402
from oracle import mfcd_oracle as O  # noqa: E402
from oracle import epoch_oracle as E  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def store_of(cols, dev):
    u, i, j, z = cols
    return TripletStore.from_columns(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(j),
                                     torch.from_numpy(z), device=dev)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if KIND == "gloo1":
        dev = torch.device("cuda", 0)
        dist.init_process_group("gloo")
        backend = "allreduce"
    else:
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
        backend = "peer"
    torch.cuda.set_device(dev)
    report = {"world": world, "kind": KIND}
    rng = np.random.default_rng(11)          # same data on every rank; each takes its shard
    n, m, d = 60, 40, 8

    def columns(N):
        return (rng.integers(0, n, N), rng.integers(0, m, N), rng.integers(0, m, N),
                rng.integers(0, 2, N).astype(np.float64))

    def shard(cols, r):
        N = len(cols[0]) // world
        return tuple(c[r * N:(r + 1) * N] for c in cols)

    # ---- 1. deterministic, in-order loaders -------------------------------------------------------------
    Ntr, Nva, Bg = 1280, 256, 128
    tr, va, te = columns(Ntr), columns(Nva), columns(512)
    torch.manual_seed(3)
    model = MatrixFactorization(n, m, d)
    U0, V0 = model.U.detach().numpy().copy(), model.V.detach().numpy().copy()
    lr, wd = 1e-2, 1e-4
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    tl = TripletLoader(store_of(shard(tr, rank), dev), Bg, shuffle=False)
    vl = TripletLoader(store_of(shard(va, rank), dev), Bg, shuffle=False)
    t_losses, v_losses = trainer.train_model(model, tl, vl, opt, dev, num_epochs=2, mode="deterministic",
                                             world_size=world, dp_backend=backend)
    Bl = Bg // world
    shards = [shard(tr, r) for r in range(world)]
    gb = []
    for k in range(Ntr // Bg):
        gb.append(tuple(np.concatenate([s[c][k * Bl:(k + 1) * Bl] for s in shards]) for c in range(4)))
    Uo, Vo = U0.copy(), V0.copy()
    l1, st = O.train_steps(Uo, Vo, gb, lr, wd)
    l2, st = O.train_steps(Uo, Vo, gb, lr, wd, state=st)
    report["det_loss_rel"] = rel(t_losses, [np.mean(l1), np.mean(l2)])
    report["det_U_rel"] = rel(model.U.detach().cpu().numpy(), Uo)
    report["det_V_rel"] = rel(model.V.detach().cpu().numpy(), Vo)
    # validation: mean over ALL ranks' local batches of the batch means
    vb = []
    for r in range(world):
        s = shard(va, r)
        vb += O.split_batches(*s, Bg)
    report["det_val_rel"] = rel(v_losses[-1], O.mean_of_batch_means(Uo, Vo, vb))
    loss, acc = trainer.evaluate_model(model, TripletLoader(store_of(shard(te, rank), dev), 64), dev, world_size=world)
    tb = []
    for r in range(world):
        tb += O.split_batches(*shard(te, r), 64)
    ol, oa = O.evaluate_model(Uo, Vo, tb)
    report["eval_loss_rel"] = rel(loss, ol)
    report["eval_acc_abs"] = abs(acc - oa)
    fs = model.flat_state(dev)
    mine_p = fs.params.detach().cpu().contiguous() if KIND == "gloo1" else fs.params.contiguous()
    gathered = [torch.zeros_like(mine_p) for _ in range(world)]
    dist.all_gather(gathered, mine_p)
    report["replicas_identical"] = all(torch.equal(gathered[0], t) for t in gathered)

    # ---- 2. atomic, throughput batch, per-epoch device reshuffle + grouping ------------------------------
    n2, m2, d2, Nl, Bg2 = 300, 200, 16, 8192, 2048
    rng2 = np.random.default_rng(100 + rank)     # this rank's own shard
    sh = (rng2.integers(0, n2, Nl), rng2.integers(0, m2, Nl), rng2.integers(0, m2, Nl),
          rng2.integers(0, 2, Nl).astype(np.float64))
    torch.manual_seed(4)
    model2 = MatrixFactorization(n2, m2, d2)
    U0, V0 = model2.U.detach().numpy().copy(), model2.V.detach().numpy().copy()
    opt2 = torch.optim.Adam(model2.parameters(), lr=lr, weight_decay=wd)
    torch.manual_seed(50 + rank)                 # ranks reshuffle independently
    tl2 = TripletLoader(store_of(sh, dev), Bg2, shuffle=True, shuffle_rng="device")
    vl2 = TripletLoader(store_of(tuple(c[:256] for c in sh), dev), Bg2, shuffle=False, shuffle_rng="device")
    t2, _ = trainer.train_model(model2, tl2, vl2, opt2, dev, num_epochs=1, mode="atomic", world_size=world,
                                dp_backend=backend)
    mine = {"seed": tl2.last_epoch_seed, "rec": tl2.store.rec.cpu().numpy()}     # store is user-sorted now
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    Bl2 = Bg2 // world
    per_rank = []
    for e in everyone:
        pos = E.epoch_positions(Nl, e["seed"])
        recs, _ = E.epoch_batches(e["rec"], pos, Bl2)
        per_rank.append(recs)
    gb2 = []
    for k in range(Nl // Bl2):
        rows = np.concatenate([pr[k * Bl2:(k + 1) * Bl2] for pr in per_rank])
        gb2.append((rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64), rows[:, 2].astype(np.int64),
                    rows[:, 3].copy().view(np.float32).astype(np.float64)))
    Uo, Vo = U0.copy(), V0.copy()
    lo, _ = O.train_steps(Uo, Vo, gb2, lr, wd)
    report["atomic_loss_rel"] = rel(t2[0], np.mean(lo))
    report["atomic_U_rel"] = rel(model2.U.detach().cpu().numpy(), Uo)
    report["atomic_V_rel"] = rel(model2.V.detach().cpu().numpy(), Vo)
    fs2 = model2.flat_state(dev)
    mine_p = fs2.params.detach().cpu().contiguous() if KIND == "gloo1" else fs2.params.contiguous()
    gathered = [torch.zeros_like(mine_p) for _ in range(world)]
    dist.all_gather(gathered, mine_p)
    report["atomic_replicas_identical"] = all(torch.equal(gathered[0], t) for t in gathered)
    report["grads_left_clean"] = bool((fs2.grads == 0).all().item())

    ok = (report["det_loss_rel"] < 1e-5 and report["det_U_rel"] < 1e-5 and report["det_V_rel"] < 1e-5
          and report["det_val_rel"] < 1e-5 and report["eval_loss_rel"] < 1e-5 and report["eval_acc_abs"] < 1e-9
          and report["replicas_identical"] and report["atomic_loss_rel"] < 2e-5 and report["atomic_U_rel"] < 2e-5
          and report["atomic_V_rel"] < 2e-5 and report["atomic_replicas_identical"] and report["grads_left_clean"])
    report["ok"] = bool(ok)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dp_api_{KIND}_w{world}.json"), "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
