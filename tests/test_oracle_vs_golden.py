"""Pin the numpy oracle against outputs of the reference (tests/golden/*.npz,
made by tests/golden/make_golden.py which runs /root/reference unmodified)."""
import math

import numpy as np
import pytest

from conftest import load_golden, batches_from
from oracle import mfcd_oracle as O

TRAIN_FIXTURES = ["train_c1.npz", "train_d10_k3.npz", "train_soft_d4.npz", "train_d64.npz"]
KAT_TAGS = ["dup", "sat", "soft", "d2", "d3", "d32", "d64", "d128", "single"]


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("tag", KAT_TAGS)
def test_kat_forward_backward(tag):
    g = load_golden("kat_fwd_bwd.npz")
    U, V = g[tag + "_U"], g[tag + "_V"]
    u, i, j, z = g[tag + "_u"], g[tag + "_i"], g[tag + "_j"], g[tag + "_z"]
    p = O.forward(U, V, u, i, j)
    assert rel(p, g[tag + "_pred"]) < 2e-6
    loss, gU, gV = O.loss_and_grads(U, V, u, i, j, z.astype(np.float32))
    assert abs(loss - g[tag + "_loss"]) <= 1e-5 * abs(g[tag + "_loss"]) + 1e-7
    assert rel(gU, g[tag + "_gU"]) < 1e-5
    assert rel(gV, g[tag + "_gV"]) < 1e-5


def test_kat_saturation_is_exercised():
    g = load_golden("kat_fwd_bwd.npz")
    p = g["sat_pred"]
    assert (p == 0.0).any() and (p == 1.0).any()      # fp32 saturation happened
    assert g["sat_loss"] > 10.0                        # -100 clamp reached


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_train_replay(name):
    g = load_golden(name)
    U, V = g["U0"].copy(), g["V0"].copy()
    batches = batches_from(g)
    losses, st = O.train_steps(U, V, batches, float(g["lr"]), float(g["wd"]))
    assert rel(losses, g["step_losses"]) < 1e-5
    assert rel(U, g["U1"]) < 1e-5 and rel(V, g["V1"]) < 1e-5
    spe = int(g["steps_per_epoch"])
    ep = [float(np.mean(losses[e * spe:(e + 1) * spe])) for e in range(int(g["epochs"]))]
    assert rel(ep, g["train_losses"]) < 1e-5


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_step_snapshots(name):
    g = load_golden(name)
    U, V = g["U0"].copy(), g["V0"].copy()
    batches = batches_from(g)
    st = None
    done = 0
    for k in (1, 2, 5):
        _, st = O.train_steps(U, V, batches[done:k], float(g["lr"]), float(g["wd"]), state=st)
        done = k
        assert rel(U, g[f"U_step{k}"]) < 1e-5 and rel(V, g[f"V_step{k}"]) < 1e-5


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_eval_metrics(name):
    g = load_golden(name)
    U, V, X = g["U1"], g["V1"], g["X"]
    test = O.split_batches(g["test_u"], g["test_i"], g["test_j"], g["test_z"], 64)
    val = O.split_batches(g["val_u"], g["val_i"], g["val_j"], g["val_z"], 64)
    loss, acc = O.evaluate_model(U, V, test)
    assert abs(loss - g["test_loss"]) < 1e-5 * abs(g["test_loss"])
    assert acc == pytest.approx(float(g["test_acc"]), abs=1e-12)
    assert abs(O.mean_of_batch_means(U, V, val) - g["val_losses"][-1]) < 1e-5
    assert abs(O.reconstruction_error(U, V, X, float(g["s"])) - g["rec_err"]) < 1e-4 * g["rec_err"]
    gl, ga = O.ground_truth_metrics(X, test)
    assert abs(gl - g["gt_loss"]) < 1e-5 and ga == pytest.approx(float(g["gt_acc"]), abs=1e-12)


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_alpha_and_norm_ratios(name):
    g = load_golden(name)
    out = O.alpha_and_norm_ratios(g["U1"], g["V1"], g["X"])
    scal = [out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7], out[8], out[12]]
    ref = g["alpha_scalars"]
    for k, (a, b) in enumerate(zip(scal, ref)):
        assert abs(a - b) <= 1e-3 * max(abs(b), 1e-3), (k, a, b)
    assert len(out[9]) == len(g["slopes"]) and rel(out[9], g["slopes"]) < 1e-3
    assert len(out[10]) == len(g["correlations"]) and np.abs(np.array(out[10]) - g["correlations"]).max() < 1e-3
    assert len(out[11]) == len(g["spearman_scores"]) and np.abs(np.array(out[11]) - g["spearman_scores"]).max() < 1e-3
    assert rel(out[13], g["alpha_per_row"]) < 1e-3


def test_split_sizes_and_topup():
    g = load_golden("samplers.npz")
    # 100 triplets, K=3, soft labels on train only: 80 soft rows, 10*3 val rows, test topped up to ceil(500/3)
    assert O.split_sizes(100) == (80, 10, 10)
    assert len(g["split_tr_u"]) == 80 and len(g["split_va_u"]) == 30
    assert O.test_topup_needed(10, 3) == 167 - 10
    assert len(g["split_te_u"]) == 167 * 3
    zs = g["split_tr_z"]
    assert np.allclose(zs * 3, np.round(zs * 3))       # soft labels are multiples of 1/K
    assert set(np.unique(g["split_te_z"])) <= {0.0, 1.0}
    # hard rows come as K consecutive copies of one triplet
    te = np.stack([g["split_te_u"], g["split_te_i"], g["split_te_j"]], 1).reshape(-1, 3, 3)
    assert (te == te[:, :1]).all()


def test_sampler_rules():
    g = load_golden("samplers.npz")
    X = g["X"]
    n, m = X.shape
    for tag in ("random", "popularity_zipf", "popularity_exp", "margin", "svd"):
        t = g[tag]
        assert (t[:, 1] != t[:, 2]).all()
        assert len({tuple(r) for r in t.tolist()}) == len(t)
        assert t[:, 0].min() >= 0 and t[:, 0].max() < n and t[:, 1:].max() < m
    thr = O.margin_threshold(X, 300)
    assert abs(thr - g["margin_value"]) < 1e-6 * abs(g["margin_value"])
    pred = O.margin_predicate(X, thr)
    assert all(pred(*r) for r in g["margin"].tolist())
    rank = O.svd_rank(n, m, int(g["svd_num"]))
    assert rank == 4
    users, items, _, _ = O.svd_top_sets(X, min(rank, 3))   # X has rank 3: extra directions carry ~0 weight
    t = g["svd"]
    assert set(t[:, 0].tolist()) <= set(users.tolist())
    assert set(t[:, 1].tolist()) | set(t[:, 2].tolist()) <= set(items.tolist())
    # popularity: low item ids dominate, as the zipf law over index order says
    pz = O.popularity_probs(m, "zipf", 1.5)
    assert pz[0] > pz[1] > pz[-1] and abs(pz.sum() - 1) < 1e-12
    cnt = np.bincount(g["popularity_zipf"][:, 1:].ravel(), minlength=m)
    assert cnt[:5].sum() > cnt[-5:].sum()
    with pytest.raises(ValueError):
        O.popularity_probs(m, "nope")


def test_accept_stream_and_labels():
    cands = [(0, 1, 1), (0, 1, 2), (0, 1, 2), (1, 2, 0), (2, 0, 1), (3, 1, 0)]
    kept, used = O.accept_stream(cands, 3, exclude={(1, 2, 0)})
    assert kept == [(0, 1, 2), (2, 0, 1), (3, 1, 0)] and used == 6
    q = np.array([0.25, 0.75], np.float32)
    uni = np.array([0.1, 0.3, 0.2, 0.9, 0.7, 0.8], np.float32)
    assert O.btl_labels_from_uniforms(q, uni, 3, soft=False).tolist() == [1, 0, 1, 0, 1, 0]
    assert np.allclose(O.btl_labels_from_uniforms(q, uni, 3, soft=True), [2 / 3, 1 / 3])


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_torch_port_replays_the_reference_bit_for_bit(name):
    """oracle/torch_port.py runs the same ATen ops as the reference: identical losses and weights."""
    import torch
    from oracle import torch_port as TP
    torch.set_num_threads(1)
    g = load_golden(name)
    model = TP.PortModel(torch.from_numpy(g["U0"]), torch.from_numpy(g["V0"]))
    opt = TP.make_optimizer(model, float(g["lr"]), float(g["wd"]))
    losses = [TP.train_step(model, opt, torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(j), torch.from_numpy(z))
              for (u, i, j, z) in batches_from(g)]
    assert np.array_equal(np.array(losses), g["step_losses"])
    assert np.array_equal(model.U.detach().numpy(), g["U1"]) and np.array_equal(model.V.detach().numpy(), g["V1"])
    test = [tuple(torch.from_numpy(a) for a in b)
            for b in O.split_batches(g["test_u"], g["test_i"], g["test_j"], g["test_z"], 64)]
    loss, acc = TP.eval_batches(model, test)
    assert loss == pytest.approx(float(g["test_loss"]), rel=1e-12) and acc == pytest.approx(float(g["test_acc"]), abs=1e-12)


def test_config2_first_2048_steps():
    """BASELINE config 2 (1000 x 1000, d = 10, K = 3): the numpy oracle replays the first 2048 optimiser steps the
    reference recorded (train_c2_steps.npz) -- per-step losses and final weights within 1e-5."""
    g = load_golden("train_c2_steps.npz")
    U, V = g["U0"].copy(), g["V0"].copy()
    steps = int(g["steps"])
    bu, bi, bj, bz = (g["batch_u"].astype(np.int64), g["batch_i"].astype(np.int64), g["batch_j"].astype(np.int64),
                      g["batch_z"].astype(np.float64))
    assert len(bu) == steps * 64 and bu.max() < 1000 and set(np.unique(bz)) <= {0.0, 1.0}
    batches = [(bu[k * 64:(k + 1) * 64], bi[k * 64:(k + 1) * 64], bj[k * 64:(k + 1) * 64], bz[k * 64:(k + 1) * 64])
               for k in range(steps)]
    losses, _ = O.train_steps(U, V, batches, float(g["lr"]), float(g["wd"]))
    assert np.abs(np.array(losses) - g["step_losses"]).max() < 1e-5 * np.abs(g["step_losses"]).max()
    assert np.abs(U - g["U_end"]).max() < 1e-5 * np.abs(g["U_end"]).max()
    assert np.abs(V - g["V_end"]).max() < 1e-5 * np.abs(g["V_end"]).max()
