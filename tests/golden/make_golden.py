"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference; nothing of it is
copied into this repo.  Outputs are small .npz files that the CPU and GPU test
suites load.  The reference ships no tests of its own, so
these files ARE the pin for the oracle and for the CUDA path.
"""
import os
import sys
import math

REF = os.environ.get("MFCD_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import numpy as np
import torch
import torch.nn.functional as F

import structure as R           # the reference (sets OMP_NUM_THREADS=4 on import)
import generation_data as RG


class Recorder:
    """Iterates a DataLoader and remembers every batch it handed out."""

    def __init__(self, loader):
        self.loader = loader
        self.batches = []

    def __iter__(self):
        for b in self.loader:
            self.batches.append([t.clone() for t in b])
            yield b

    def __len__(self):
        return len(self.loader)


def dataset_arrays(loader):
    data = loader.dataset.data
    u = np.array([t[0] for t in data], np.int64)
    i = np.array([t[1] for t in data], np.int64)
    j = np.array([t[2] for t in data], np.int64)
    z = np.array([t[3] for t in data], np.float64)
    return u, i, j, z


def pack_batches(batches):
    u = torch.cat([b[0] for b in batches]).numpy().astype(np.int64)
    i = torch.cat([b[1] for b in batches]).numpy().astype(np.int64)
    j = torch.cat([b[2] for b in batches]).numpy().astype(np.int64)
    z = torch.cat([b[3] for b in batches]).numpy().astype(np.float64)
    sizes = np.array([len(b[0]) for b in batches], np.int64)
    return u, i, j, z, sizes


def make_train_fixture(name, n, m, d, p, s, K, soft_label, epochs, seed, strategy="random",
                       lr=1e-3, wd=1e-5):
    torch.manual_seed(seed)
    np.random.seed(seed)
    X = R.generate_X(n, m, d, "cpu")
    num_triplets = int(n * m * p / 2)
    train_loader, val_loader, test_loader = R.split_dataset_from_triplets(
        X, num_triplets, scale=s, K=K, strategy=strategy, soft_label=soft_label)
    model = R.MatrixFactorization(n, m, d)
    U0 = model.U.detach().clone().numpy()
    V0 = model.V.detach().clone().numpy()
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    rec_train = Recorder(train_loader)
    train_losses, val_losses = R.train_model(model, rec_train, val_loader, opt, "cpu",
                                             num_epochs=epochs)
    U1 = model.U.detach().clone().numpy()
    V1 = model.V.detach().clone().numpy()

    # step-by-step replay with the reference's own model class on the recorded
    # batches: gives per-step losses; must land on the same weights bit for bit.
    replay = R.MatrixFactorization(n, m, d)
    with torch.no_grad():
        replay.U.copy_(torch.from_numpy(U0))
        replay.V.copy_(torch.from_numpy(V0))
    ropt = torch.optim.Adam(replay.parameters(), lr=lr, weight_decay=wd)
    step_losses = []
    snap_steps = [1, 2, 5]
    snaps = {}
    for k, (u, i, j, z) in enumerate(rec_train.batches, 1):
        ropt.zero_grad()
        loss = F.binary_cross_entropy(replay(u, i, j), z.float())
        loss.backward()
        ropt.step()
        step_losses.append(loss.item())
        if k in snap_steps:
            snaps[k] = (replay.U.detach().clone().numpy(), replay.V.detach().clone().numpy())
    assert np.array_equal(replay.U.detach().numpy(), U1), "replay diverged from train_model"
    assert np.array_equal(replay.V.detach().numpy(), V1)

    test_loss, test_acc = R.evaluate_model(model, test_loader, "cpu")
    rec_err = R.compute_reconstruction_error(model, X, s)
    tup = R.compute_alpha_and_norm_ratios(model, X)
    gt_loss, gt_acc = R.compute_ground_truth_metrics(test_loader, X, "cpu")

    bu, bi, bj, bz, bsizes = pack_batches(rec_train.batches)
    tu, ti, tj, tz = dataset_arrays(train_loader)
    vu, vi, vj, vz = dataset_arrays(val_loader)
    eu, ei, ej, ez = dataset_arrays(test_loader)
    out = dict(
        n=n, m=m, d=d, p=p, s=s, K=K, soft_label=int(soft_label), epochs=epochs, seed=seed,
        lr=lr, wd=wd, X=X.numpy(), U0=U0, V0=V0, U1=U1, V1=V1,
        train_u=tu, train_i=ti, train_j=tj, train_z=tz,
        val_u=vu, val_i=vi, val_j=vj, val_z=vz,
        test_u=eu, test_i=ei, test_j=ej, test_z=ez,
        batch_u=bu, batch_i=bi, batch_j=bj, batch_z=bz, batch_sizes=bsizes,
        steps_per_epoch=len(train_loader),
        step_losses=np.array(step_losses, np.float64),
        train_losses=np.array(train_losses, np.float64),
        val_losses=np.array(val_losses, np.float64),
        test_loss=test_loss, test_acc=test_acc, rec_err=rec_err,
        gt_loss=gt_loss, gt_acc=gt_acc,
        alpha_scalars=np.array([tup[0], tup[1], tup[2], tup[3], tup[4], tup[5], tup[6],
                                tup[7], tup[8], tup[12]], np.float64),
        slopes=np.array(tup[9], np.float64), correlations=np.array(tup[10], np.float64),
        spearman_scores=np.array(tup[11], np.float64),
        alpha_per_row=np.array(tup[13], np.float64),
    )
    for k, (Us, Vs) in snaps.items():
        out[f"U_step{k}"] = Us
        out[f"V_step{k}"] = Vs
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "steps", len(step_losses), "train", train_losses, "acc", test_acc)


def make_kat():
    """Known-answer cases for one forward+backward via torch autograd on the
    reference MatrixFactorization: saturation, duplicates, soft labels, ragged."""
    g = torch.Generator().manual_seed(123)
    cases = {}

    def run(tag, n, m, d, u, i, j, z, scale_U=1.0):
        model = R.MatrixFactorization(n, m, d)
        with torch.no_grad():
            model.U.copy_(torch.randn(n, d, generator=g) * scale_U)
            model.V.copy_(torch.randn(m, d, generator=g) * scale_U)
        u = torch.as_tensor(u, dtype=torch.int64)
        i = torch.as_tensor(i, dtype=torch.int64)
        j = torch.as_tensor(j, dtype=torch.int64)
        z = torch.as_tensor(z, dtype=torch.float64)
        pred = model(u, i, j)
        loss = F.binary_cross_entropy(pred, z.float())
        loss.backward()
        cases[tag + "_U"] = model.U.detach().numpy().copy()
        cases[tag + "_V"] = model.V.detach().numpy().copy()
        cases[tag + "_u"] = u.numpy()
        cases[tag + "_i"] = i.numpy()
        cases[tag + "_j"] = j.numpy()
        cases[tag + "_z"] = z.numpy()
        cases[tag + "_pred"] = pred.detach().numpy().copy()
        cases[tag + "_loss"] = np.float64(loss.item())
        cases[tag + "_gU"] = model.U.grad.numpy().copy()
        cases[tag + "_gV"] = model.V.grad.numpy().copy()

    B = 64
    # plain batch, many duplicate users/items (small tables)
    run("dup", 7, 5, 4, torch.randint(0, 7, (B,), generator=g), torch.randint(0, 5, (B,), generator=g),
        torch.randint(0, 5, (B,), generator=g), torch.randint(0, 2, (B,), generator=g).double())
    # saturation: huge embeddings -> |x| >> 17, p == 0.0 or 1.0 exactly, loss clamps at 100
    run("sat", 9, 11, 8, torch.randint(0, 9, (B,), generator=g), torch.randint(0, 11, (B,), generator=g),
        torch.randint(0, 11, (B,), generator=g), torch.randint(0, 2, (B,), generator=g).double(),
        scale_U=6.0)
    # soft labels (multiples of 1/K) and a ragged last batch (B=37), d=10 (not /4)
    run("soft", 20, 30, 10, torch.randint(0, 20, (37,), generator=g), torch.randint(0, 30, (37,), generator=g),
        torch.randint(0, 30, (37,), generator=g), torch.randint(0, 4, (37,), generator=g).double() / 3.0)
    # wide rows: d=64, d=128, d=2, odd d=3, i == j rows (zero gradient contributions)
    for d in (2, 3, 32, 64, 128):
        run(f"d{d}", 33, 29, d, torch.randint(0, 33, (B,), generator=g), torch.randint(0, 29, (B,), generator=g),
            torch.randint(0, 29, (B,), generator=g), torch.randint(0, 2, (B,), generator=g).double())
    run("single", 3, 3, 2, [1], [0], [2], [1.0])
    np.savez_compressed(os.path.join(HERE, "kat_fwd_bwd.npz"), **cases)
    print("kat_fwd_bwd.npz", len(cases), "arrays")


def make_samplers():
    out = {}
    torch.manual_seed(11)
    np.random.seed(11)
    X = R.generate_X(40, 30, 3, "cpu")
    out["X"] = X.numpy()

    torch.manual_seed(3); np.random.seed(3)
    t = RG.choose_items_random(X, 200, set())
    out["random"] = np.array(t, np.int64)

    torch.manual_seed(4); np.random.seed(4)
    t = RG.choose_items_by_popularity(X, 150, set(), method="zipf", alpha=1.5)
    out["popularity_zipf"] = np.array(t, np.int64)
    torch.manual_seed(4); np.random.seed(4)
    t = RG.choose_items_by_popularity(X, 100, set(), method="exponential", alpha=0.2)
    out["popularity_exp"] = np.array(t, np.int64)

    # margin / svd call an unseeded default_rng(): pin it for the recording
    real = np.random.default_rng
    np.random.default_rng = lambda *a, **k: real(2024)
    try:
        num = 300
        t = RG.choose_items_by_margin(X, num, set())
        out["margin"] = np.array([(int(a), int(b), int(c)) for a, b, c in t], np.int64)
        sample = X[:10].numpy()
        out["margin_value"] = np.float64(np.mean(np.max(sample, axis=1) - np.min(sample, axis=1)) * num / (40 * 30))
        num = 120   # rank = int(120/1200*40) = 4 > rank(X)=3 is fine for svds (k < min(n,m))
        t = RG.choose_items_by_svd_projection(X, num, set())
        out["svd"] = np.array([(int(a), int(b), int(c)) for a, b, c in t], np.int64)
        out["svd_num"] = num
    finally:
        np.random.default_rng = real

    # split sizes + top-up + label layout under a fixed seed
    torch.manual_seed(21); np.random.seed(21)
    tr, va, te = R.split_dataset_from_triplets(X, 100, scale=2.0, K=3, soft_label=True)
    for tag, ld in (("tr", tr), ("va", va), ("te", te)):
        u, i, j, z = dataset_arrays(ld)
        out[f"split_{tag}_u"], out[f"split_{tag}_i"], out[f"split_{tag}_j"], out[f"split_{tag}_z"] = u, i, j, z
    np.savez_compressed(os.path.join(HERE, "samplers.npz"), **out)
    print("samplers.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def make_samplers_next():
    """The next-ring strategies (proximity / top_k / variance, generation_data.py:29-43, 189-224, 87-99) under
    pinned torch + numpy seeds, on the same X as samplers.npz."""
    out = {}
    torch.manual_seed(11)
    np.random.seed(11)
    X = R.generate_X(40, 30, 3, "cpu")
    out["X"] = X.numpy()
    torch.manual_seed(31); np.random.seed(31)
    out["proximity_k5"] = np.array(RG.choose_items_by_proximity(X, 200, set(), k=5), np.int64)
    torch.manual_seed(32); np.random.seed(32)
    out["top_k_default"] = np.array(RG.choose_items_top_k(X, 150, set()), np.int64)       # k = max(5, int(0.1 m)) = 5
    torch.manual_seed(33); np.random.seed(33)
    out["variance"] = np.array(RG.choose_items_by_variance(X, 400, set()), np.int64)
    out["variance_probs"] = (torch.var(X, dim=0) / torch.var(X, dim=0).sum()).numpy().astype(np.float64)
    # top_k gives up after 3 x num_triplets attempts: ask for more than the block can hold (40 users x 5 x 4 pairs = 800)
    torch.manual_seed(34); np.random.seed(34)
    out["top_k_saturated"] = np.array(RG.choose_items_top_k(X, 790, set()), np.int64)
    np.savez_compressed(os.path.join(HERE, "samplers_next.npz"), **out)
    print("samplers_next.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def make_c2_steps(steps=2048):
    """Config 2 of BASELINE.json (1000 x 1000, d = 10, p = 0.5, K = 3, random sampling): the first `steps` optimiser
    steps of the reference's first epoch -- the batches its DataLoader handed out, the per-step losses and the
    weights after the last recorded step.  Sized for the persistent small-batch kernel with several CTAs."""
    n, m, d, p, s, K, seed, lr, wd = 1000, 1000, 10, 0.5, 1.0, 3, 7, 1e-3, 1e-5
    torch.manual_seed(seed)
    np.random.seed(seed)
    X = R.generate_X(n, m, d, "cpu")
    train_loader, val_loader, test_loader = R.split_dataset_from_triplets(X, int(n * m * p / 2), scale=s, K=K)
    model = R.MatrixFactorization(n, m, d)
    U0 = model.U.detach().clone().numpy()
    V0 = model.V.detach().clone().numpy()
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    losses, bu, bi, bj, bz = [], [], [], [], []
    model.train()
    for k, (u, i, j, z) in enumerate(train_loader):           # the loop body of structure.py:845-852
        if k >= steps:
            break
        opt.zero_grad()
        loss = F.binary_cross_entropy(model(u, i, j), z.float())
        loss.backward()
        opt.step()
        losses.append(loss.item())
        bu.append(u.numpy().astype(np.int16)); bi.append(i.numpy().astype(np.int16)); bj.append(j.numpy().astype(np.int16))
        bz.append(z.numpy().astype(np.uint8))
    out = dict(n=n, m=m, d=d, p=p, s=s, K=K, seed=seed, lr=lr, wd=wd, steps=len(losses), U0=U0, V0=V0,
               U_end=model.U.detach().numpy().copy(), V_end=model.V.detach().numpy().copy(),
               batch_u=np.concatenate(bu), batch_i=np.concatenate(bi), batch_j=np.concatenate(bj),
               batch_z=np.concatenate(bz), step_losses=np.array(losses, np.float64),
               n_train_samples=len(train_loader.dataset))
    np.savez_compressed(os.path.join(HERE, "train_c2_steps.npz"), **out)
    print("train_c2_steps.npz", len(losses), "steps, first/last loss", losses[0], losses[-1])


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_train_fixture("train_c1.npz", 100, 100, 2, 0.1, 1.0, 1, False, epochs=3, seed=0)
    make_train_fixture("train_d10_k3.npz", 60, 50, 10, 0.5, 1.0, 3, False, epochs=2, seed=1)
    make_train_fixture("train_soft_d4.npz", 48, 64, 4, 0.4, 5.0, 4, True, epochs=2, seed=2)
    make_train_fixture("train_d64.npz", 96, 80, 64, 0.2, 1.0, 1, False, epochs=1, seed=3)
    make_kat()
    make_samplers()
    make_samplers_next()
    make_c2_steps()
