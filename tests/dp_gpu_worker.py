"""Data-parallel equivalence check on real GPUs; run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dp_gpu_worker.py

1. the reference's recorded first epoch (golden train_d10_k3) trained data-parallel with the SAME global
   batches (ReplicatedPlan, deterministic scatter) must reproduce the reference's per-step losses and
   weights to 1e-5, and all replicas must be bit-identical;
2. a larger synthetic problem: W-rank DP (atomic scatter, hot-row privatisation) == one GPU with batch B.
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

import mfcd_b200  # noqa: E402
from mfcd_b200 import dist as mdist  # noqa: E402
from mfcd_b200.store import TripletStore  # noqa: E402
from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch  # noqa: E402
from conftest import load_golden  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    rank, world = mdist.init_from_env("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    report = {"world": world}

    # ---- 1. golden epoch, deterministic, replicated plan ------------------------------------------------
    g = load_golden("train_d10_k3.npz")
    spe = int(g["steps_per_epoch"])
    N = int(g["batch_sizes"][:spe].sum())
    store = TripletStore.from_columns(torch.from_numpy(g["batch_u"][:N]), torch.from_numpy(g["batch_i"][:N]),
                                      torch.from_numpy(g["batch_j"][:N]), torch.from_numpy(g["batch_z"][:N]), device=dev)
    model = MatrixFactorization(int(g["n"]), int(g["m"]), int(g["d"]))
    with torch.no_grad():
        model.U.copy_(torch.from_numpy(g["U0"])); model.V.copy_(torch.from_numpy(g["V0"]))
    fs = model.flat_state(dev)
    spec = OptimizerSpec.adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    eng = mdist.CudaEngine(fs, store, None, spec, 1)
    plan = mdist.ReplicatedPlan(N, 64, rank, world)
    losses = torch.zeros(plan.n_steps(), dtype=torch.float32, device=dev)
    mdist.dp_epoch(eng, plan, 0, losses, bucket_elems=256)
    torch.cuda.synchronize()
    report["golden_loss_rel"] = rel(losses.cpu().numpy(), g["step_losses"][:spe])
    # weights after epoch 1 are not stored in the golden file: compare with a single-GPU replay of the same epoch
    ref_model = MatrixFactorization(int(g["n"]), int(g["m"]), int(g["d"]))
    with torch.no_grad():
        ref_model.U.copy_(torch.from_numpy(g["U0"])); ref_model.V.copy_(torch.from_numpy(g["V0"]))
    rfs = ref_model.flat_state(dev)
    rl = run_epoch(rfs, store, None, 64, spec, 1)
    report["vs_single_gpu_params_rel"] = rel(fs.params.cpu().numpy(), rfs.params.cpu().numpy())
    report["vs_single_gpu_loss_rel"] = rel(losses.cpu().numpy(), rl.cpu().numpy())
    gathered = [torch.zeros_like(fs.params) for _ in range(world)]
    dist.all_gather(gathered, fs.params)
    report["replicas_identical"] = all(torch.equal(gathered[0], t) for t in gathered)

    # ---- 2. synthetic, atomic + hot rows, partitioned-free comparison against one GPU ---------------------
    n, m, d, B, steps = 5000, 3000, 64, 1 << 16, 6
    gen = torch.Generator(device=dev); gen.manual_seed(77)            # same data on every rank
    Ntot = B * steps
    rec = torch.empty((Ntot, 4), dtype=torch.int32, device=dev)
    rec[:, 0] = torch.randint(0, n, (Ntot,), generator=gen, device=dev)
    pr = 1.0 / torch.arange(1, m + 1, device=dev, dtype=torch.float64) ** 1.5
    rec[:, 1] = torch.multinomial(pr, Ntot, replacement=True, generator=gen)
    rec[:, 2] = (rec[:, 1] + 1 + torch.multinomial(pr, Ntot, replacement=True, generator=gen)) % m
    rec[:, 3] = torch.randint(0, 2, (Ntot,), generator=gen, device=dev).float().view(torch.int32)
    big = TripletStore(rec)
    torch.manual_seed(5)
    m_dp = MatrixFactorization(n, m, d); m_one = MatrixFactorization(n, m, d)
    with torch.no_grad():
        m_one.U.copy_(m_dp.U); m_one.V.copy_(m_dp.V)
    f_dp, f_one = m_dp.flat_state(dev), m_one.flat_state(dev)
    spec2 = OptimizerSpec.adam(lr=1e-2, weight_decay=1e-5)
    eng2 = mdist.CudaEngine(f_dp, big, None, spec2, 0)
    plan2 = mdist.ReplicatedPlan(Ntot, B, rank, world)
    l_dp = torch.zeros(steps, dtype=torch.float32, device=dev)
    mdist.dp_epoch(eng2, plan2, 0, l_dp)
    l_one = run_epoch(f_one, big, None, B, spec2, 0)
    torch.cuda.synchronize()
    report["synthetic_loss_rel"] = rel(l_dp.cpu().numpy(), l_one.cpu().numpy())
    report["synthetic_params_rel"] = rel(f_dp.params.cpu().numpy(), f_one.params.cpu().numpy())
    # ---- 3. fused peer-memory exchange (K9) == NCCL all-reduce + K3 --------------------------------------
    peer_ok = True
    for tag, mm in (("peer_p2p", False), ("peer_multimem", True)):
        try:
            ex = mdist.PeerExchange((n + m) * d, dev, use_multimem=mm)
            if mm and not ex.multimem:
                report[tag] = "multicast not available on this box"
                continue
            torch.manual_seed(5)
            m_px = MatrixFactorization(n, m, d)               # same init stream as m_dp / m_one
            f_px = m_px.flat_state(dev, storage=ex.storage())
            eng3 = mdist.CudaEngine(f_px, big, None, spec2, 0)
            l_px = torch.zeros(steps, dtype=torch.float32, device=dev)
            mdist.dp_epoch(eng3, plan2, 0, l_px, exchange=ex)
            torch.cuda.synchronize()
            report[tag + "_loss_rel"] = rel(l_px.cpu().numpy(), l_dp.cpu().numpy())
            report[tag + "_params_rel"] = rel(f_px.params.cpu().numpy(), f_dp.params.cpu().numpy())
            gathered = [torch.zeros_like(f_px.params) for _ in range(world)]
            dist.all_gather(gathered, f_px.params.contiguous())
            report[tag + "_replicas_identical"] = all(torch.equal(gathered[0], t) for t in gathered)
            peer_ok = peer_ok and report[tag + "_loss_rel"] < 1e-4 and report[tag + "_params_rel"] < 2e-3 \
                and report[tag + "_replicas_identical"]
        except Exception as e:     # report, do not hide: the NCCL path above is still validated
            import traceback
            report[tag] = "FAILED: " + repr(e) + " | " + traceback.format_exc()[-600:]
            peer_ok = False
    report["peer_ok"] = bool(peer_ok)
    ok = (report["golden_loss_rel"] < 1e-5 and report["vs_single_gpu_params_rel"] < 1e-5 and report["replicas_identical"]
          and report["vs_single_gpu_loss_rel"] < 1e-5 and report["synthetic_loss_rel"] < 1e-4
          and report["synthetic_params_rel"] < 2e-3 and report["peer_ok"])
    report["ok"] = bool(ok)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dp_check_w{world}.json"), "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
