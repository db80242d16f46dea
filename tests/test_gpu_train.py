"""GPU parity tests for the training hot path (K1, K2, K3, epoch runner), all
through the C ABI, against the golden fixtures recorded from the reference and
against the numpy oracle.  Tolerances: 1e-5 relative for per-step loss and
updated U/V in deterministic mode (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden, batches_from
from oracle import mfcd_oracle as O

pytestmark = pytest.mark.gpu

KAT_TAGS = ["dup", "sat", "soft", "d2", "d3", "d32", "d64", "d128", "single"]
TRAIN_FIXTURES = ["train_c1.npz", "train_d10_k3.npz", "train_soft_d4.npz", "train_d64.npz"]


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("mode", ["atomic", "deterministic"])
@pytest.mark.parametrize("tag", KAT_TAGS)
def test_kat_forward_backward(G, tag, mode):
    g = load_golden("kat_fwd_bwd.npz")
    U, V = g[tag + "_U"], g[tag + "_V"]
    store = G.store_from(g[tag + "_u"], g[tag + "_i"], g[tag + "_j"], g[tag + "_z"])
    loss, gU, gV = G.fwd_bwd(U, V, store, mode=mode)
    assert abs(loss - g[tag + "_loss"]) <= 1e-5 * abs(g[tag + "_loss"]) + 1e-7
    assert G.rel(gU, g[tag + "_gU"]) < 1e-5
    assert G.rel(gV, g[tag + "_gV"]) < 1e-5


def test_scores_match_reference_forward(G):
    from mfcd_b200.trainer import MatrixFactorization
    g = load_golden("kat_fwd_bwd.npz")
    for tag in ("sat", "soft", "d64", "d3"):
        U, V = g[tag + "_U"], g[tag + "_V"]
        model = MatrixFactorization(U.shape[0], V.shape[0], U.shape[1])
        with torch.no_grad():
            model.U.copy_(torch.from_numpy(U)); model.V.copy_(torch.from_numpy(V))
        p = model(torch.from_numpy(g[tag + "_u"]), torch.from_numpy(g[tag + "_i"]), torch.from_numpy(g[tag + "_j"]))
        assert G.rel(p.cpu().numpy(), g[tag + "_pred"]) < 2e-6
    p_sat = g["sat_pred"]
    assert (p_sat == 0).any() and (p_sat == 1).any()


def test_empty_and_ragged_batches(G):
    g = load_golden("kat_fwd_bwd.npz")
    U, V = g["soft_U"], g["soft_V"]
    store = G.store_from(g["soft_u"], g["soft_i"], g["soft_j"], g["soft_z"])
    loss, gU, gV = G.fwd_bwd(U, V, store, B=0)
    assert loss == 0.0 and not gU.any() and not gV.any()
    # a sub-range [5, 5+13) equals the oracle on that slice
    sl = slice(5, 18)
    lo, gUo, gVo = O.loss_and_grads(U, V, g["soft_u"][sl], g["soft_i"][sl], g["soft_j"][sl], g["soft_z"][sl].astype(np.float32))
    for mode in ("atomic", "deterministic"):
        loss, gU, gV = G.fwd_bwd(U, V, store, start=5, B=13, mode=mode)
        assert abs(loss - lo) < 1e-5 * abs(lo) and G.rel(gU, gUo) < 1e-5 and G.rel(gV, gVo) < 1e-5


def _random_problem(rng, n, m, d, B, hot=False):
    U = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    V = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    u = rng.integers(0, n, B)
    if hot:   # zipf-like: a few very hot items -> long segments that straddle many chunks
        pr = 1.0 / np.arange(1, m + 1) ** 1.5; pr /= pr.sum()
        i = rng.choice(m, B, p=pr); j = rng.choice(m, B, p=pr)
    else:
        i = rng.integers(0, m, B); j = rng.integers(0, m, B)
    z = rng.integers(0, 2, B).astype(np.float64)
    return U, V, u, i, j, z


@pytest.mark.parametrize("d", [2, 10, 64, 128, 256])
@pytest.mark.parametrize("hot", [False, True])
def test_large_batch_paths_against_oracle(G, d, hot):
    """B > 256 takes the sort + segmented-reduction path; permuted access; both modes vs oracle."""
    rng = np.random.default_rng(100 + d + hot)
    n, m, B = 300, 200, 3000
    U, V, u, i, j, z = _random_problem(rng, n, m, d, B, hot)
    perm = rng.permutation(B)
    store = G.store_from(u, i, j, z)
    lo, gUo, gVo = O.loss_and_grads(U, V, u[perm], i[perm], j[perm], z[perm].astype(np.float32))
    # deterministic engines (fixed-point integer atomics = the default, and the sort + segmented reduction): the
    # north-star bar of 1e-5; atomic fp32 reductions: 2e-5
    for mode, tol in (("atomic", 2e-5), ("deterministic", 1e-5), ("deterministic_sort", 1e-5), ("deterministic_fixed", 1e-5)):
        loss, gU, gV = G.fwd_bwd(U, V, store, mode=mode, perm=perm)
        assert abs(loss - lo) < tol * abs(lo), (mode, loss, lo)
        assert G.rel(gU, gUo) < tol and G.rel(gV, gVo) < tol, mode


@pytest.mark.parametrize("d", [32, 64, 128, 256])
def test_hot_row_privatisation_matches_oracle(G, d):
    """zipf items: the shared-memory privatised atomic kernel == oracle == plain atomic kernel."""
    import ctypes as C
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(d)
    n, m, B = 400, 300, 20000
    U, V, u, i, j, z = _random_problem(rng, n, m, d, B, hot=True)
    keep = i != j
    u, i, j, z = u[keep], i[keep], j[keep], z[keep]
    B = len(u)
    store = G.store_from(u, i, j, z)
    hot = store.hot_items(m, d, B, min_hits_per_batch=200)
    assert hot is not None and 0 < hot[1].numel() <= 127 and int(hot[0][hot[1][0].item()]) == 0
    cap = C.c_int32(0)
    check(lib.mfcd_max_hot_items(d, C.byref(cap)), "cap")
    assert hot[1].numel() <= cap.value
    lo, gUo, gVo = O.loss_and_grads(U, V, u, i, j, z.astype(np.float32))
    Ud, Vd = G.dev_f32(U), G.dev_f32(V)
    gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd); loss = torch.zeros(1, device=G.DEV)
    check(lib.mfcd_triplet_fwd_bwd_hot(ptr(Ud), ptr(Vd), ptr(store.rec), None, 0, B, d, 1.0 / B, ptr(gU), ptr(gV),
                                       ptr(loss), ptr(hot[0]), ptr(hot[1]), hot[1].numel(), current_stream()), "hot")
    assert abs(loss.item() - lo) < 2e-5 * abs(lo)
    assert G.rel(gU.cpu().numpy(), gUo) < 2e-5 and G.rel(gV.cpu().numpy(), gVo) < 2e-5
    _, gU2, gV2 = G.fwd_bwd(U, V, store, mode="atomic")
    assert G.rel(gV.cpu().numpy(), gV2) < 2e-5
    # too many hot rows for the shared-memory budget is an argument error, not a crash
    rc = lib.mfcd_triplet_fwd_bwd_hot(ptr(Ud), ptr(Vd), ptr(store.rec), None, 0, B, d, 1.0 / B, ptr(gU), ptr(gV),
                                      ptr(loss), ptr(hot[0]), ptr(hot[1]), cap.value + 1, current_stream())
    assert rc == -1


def test_group_by_user_keeps_batches_and_brings_users_together(G):
    """mfcd_group_by_user: every batch holds the same records as before, one user's triplets are adjacent,
    and a user's triplets keep their order (stable); ragged last batch included."""
    rng = np.random.default_rng(3)
    N, B, n, m = 10_000, 768, 97, 50
    u, i, j = rng.integers(0, n, N), rng.integers(0, m, N), rng.integers(0, m, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    store = G.store_from(u, i, j, z)
    before = store.rec.cpu().numpy().copy()
    store.group_by_user(B)
    after = store.rec.cpu().numpy()
    assert store.k1_flags(B) == 1 and store.k1_flags(B + 1) == 0 and store.k1_flags(B, perm=object()) == 0
    for s0 in range(0, N, B):
        a, b = before[s0:s0 + B], after[s0:s0 + B]
        order = np.argsort(a[:, 0], kind="stable")
        assert np.array_equal(a[order], b)              # same records, sorted by user, stable


@pytest.mark.parametrize("N,n", [(1, 5), (31, 3), (32, 40), (8191, 70), (8192, 1), (8193, 5000), (100_003, 999)])
def test_run_length_wire_format_round_trip(G, N, n):
    """pack_wire -> (host) -> from_wire gives back the grouped batch bit for bit; sizes around word / block edges."""
    rng = np.random.default_rng(N)
    u, i, j = rng.integers(0, n, N), rng.integers(0, 65536, N), rng.integers(0, 65536, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    store = G.store_from(u, i, j, z).group_by_user(N)
    wire = store.pack_wire(0, N)
    runs = len(np.unique(u))
    assert int(wire[0].item()) == runs and int(wire[1].item()) == N
    assert wire.numel() == 4 + (N + 31) // 32 * 3 + N + runs
    back = G.TripletStore.from_wire(wire.cpu().to(G.DEV), N)          # through host memory, like the bench
    assert torch.equal(back.rec, store.rec)
    # the device packer against the numpy restatement of the layout, bit for bit, and the device unpacker on
    # words the oracle packed
    from oracle import wire_oracle as W
    gu, gi, gj, gz = [c.cpu().numpy() for c in store.columns()]
    ow = W.pack_wire(gu, gi, gj, gz)
    assert np.array_equal(wire.cpu().numpy().view(np.uint32), ow)
    back2 = G.TripletStore.from_wire(torch.from_numpy(ow.view(np.int32)).to(G.DEV), N)
    assert torch.equal(back2.rec, store.rec)
    # soft labels / wide item ids are refused, not mangled
    bad = G.store_from([0], [70000], [1], [1.0])
    with pytest.raises(Exception):
        bad.pack_wire(0, 1)
    soft = G.store_from([0], [2], [1], [0.5])
    with pytest.raises(Exception):
        soft.pack_wire(0, 1)


@pytest.mark.parametrize("d", [8, 16, 32, 64])
@pytest.mark.parametrize("flags", [0, 1])
def test_hot_row_every_lane_group_hits_the_same_row(G, d, flags):
    """Worst case for the per-warp shared hot image: in EVERY round all lane groups of a warp update the
    same privatised row (item 0 as i, item 1 as j in every triplet), so a lost update between the groups'
    turns would show up as a wrong gV[0] / gV[1]."""
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(7 + d)
    n, m, B = 97, 40, 8192 + 37
    U = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    V = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    u = np.sort(rng.integers(0, n, B))
    i = np.zeros(B, np.int64)
    j = np.ones(B, np.int64)
    z = rng.integers(0, 2, B).astype(np.float64)
    store = G.store_from(u, i, j, z)
    slot = torch.full((m,), -1, dtype=torch.int8, device=G.DEV)
    slot[0], slot[1], slot[5] = 0, 1, 2
    items = torch.tensor([0, 1, 5], dtype=torch.int32, device=G.DEV)
    lo, gUo, gVo = O.loss_and_grads(U, V, u, i, j, z.astype(np.float32))
    Ud, Vd = G.dev_f32(U), G.dev_f32(V)
    gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd); loss = torch.zeros(1, device=G.DEV)
    check(lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(store.rec), None, 0, B, d, 1.0 / B, ptr(gU), ptr(gV),
                                      ptr(loss), ptr(slot), ptr(items), 3, flags, current_stream()), "ex")
    assert abs(loss.item() - lo) < 2e-5 * abs(lo)
    assert G.rel(gU.cpu().numpy(), gUo) < 2e-5 and G.rel(gV.cpu().numpy(), gVo) < 2e-5


@pytest.mark.parametrize("d,hot", [(64, False), (64, True), (32, True), (128, True), (256, False), (16, False)])
def test_user_grouped_kernel_matches_oracle(G, d, hot):
    """K1 with MFCD_FLAG_USER_GROUPED (one U read / one gU reduction per run of equal users, hottest item
    rows in registers) == oracle, on grouped batches, on a ragged tail, and on UNGROUPED data (the flag is
    only a hint)."""
    import ctypes as C
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(100 + d + hot)
    n, m, B = 150, 300, 20011
    U, V, u, i, j, z = _random_problem(rng, n, m, d, B, hot=hot)
    keep = i != j
    u, i, j, z = u[keep], i[keep], j[keep], z[keep]
    B = len(u)
    lo, gUo, gVo = O.loss_and_grads(U, V, u, i, j, z.astype(np.float32))
    Ud, Vd = G.dev_f32(U), G.dev_f32(V)
    for grouped in (True, False):
        store = G.store_from(u, i, j, z)
        if grouped:
            store.group_by_user(B)
        hi = store.hot_items(m, d, B, min_hits_per_batch=200) if hot else None
        slot, items, nh = (ptr(hi[0]), ptr(hi[1]), hi[1].numel()) if hi else (None, None, 0)
        gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd); loss = torch.zeros(1, device=G.DEV)
        check(lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(store.rec), None, 0, B, d, 1.0 / B, ptr(gU), ptr(gV),
                                          ptr(loss), slot, items, nh, 1, current_stream()), "ex")
        assert abs(loss.item() - lo) < 2e-5 * abs(lo), (grouped, loss.item(), lo)
        assert G.rel(gU.cpu().numpy(), gUo) < 2e-5 and G.rel(gV.cpu().numpy(), gVo) < 2e-5, grouped
        if grouped and m <= 65536:
            # the same batch in the run-length wire format, decoded by K1 itself (MFCD_FLAG_WIRE_RLE)
            wire = store.pack_wire(0, B)
            gU2 = torch.zeros_like(Ud); gV2 = torch.zeros_like(Vd); loss2 = torch.zeros(1, device=G.DEV)
            check(lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(wire), None, 0, B, d, 1.0 / B, ptr(gU2), ptr(gV2),
                                              ptr(loss2), slot, items, nh, 1 | 2, current_stream()), "ex-wire")
            assert abs(loss2.item() - lo) < 2e-5 * abs(lo)
            assert G.rel(gU2.cpu().numpy(), gUo) < 2e-5 and G.rel(gV2.cpu().numpy(), gVo) < 2e-5
            rc = lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(wire), None, 5, B - 5, d, 1.0 / B, ptr(gU2), ptr(gV2),
                                             ptr(loss2), slot, items, nh, 1 | 2, current_stream())
            assert rc == -1                                  # a wire batch is read whole
    # two half batches (second one starts mid-run and ends on a partial tile) accumulate to the full batch
    store = G.store_from(u, i, j, z).group_by_user(B)
    gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd); loss = torch.zeros(1, device=G.DEV)
    h = B // 2 + 5
    for s0, cnt in ((0, h), (h, B - h)):
        check(lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(store.rec), None, s0, cnt, d, 1.0 / B, ptr(gU), ptr(gV),
                                          ptr(loss), None, None, 0, 1, current_stream()), "ex")
    assert abs(loss.item() - lo) < 2e-5 * abs(lo)
    assert G.rel(gU.cpu().numpy(), gUo) < 2e-5 and G.rel(gV.cpu().numpy(), gVo) < 2e-5


@pytest.mark.parametrize("d", [64, 128, 256])
@pytest.mark.parametrize("B", [1, 15, 16, 17, 33, 257, 5000])
def test_span_kernel_small_and_ragged_batches(G, d, B):
    """K1 span (user-grouped batches, d >= 64: contiguous spans per lane group, hot-row weights as scalars): batches
    smaller than a tile / a span, ragged tails, a start offset, every item hot, and one user for the whole batch."""
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(1000 * d + B)
    n, m = 23, 40
    U = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    V = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    for one_user in (False, True):
        u = np.zeros(B + 3, np.int64) + 7 if one_user else np.sort(rng.integers(0, n, B + 3))
        i = rng.integers(0, m, B + 3); j = (i + 1 + rng.integers(0, m - 1, B + 3)) % m
        z = rng.integers(0, 2, B + 3).astype(np.float64)
        store = G.store_from(u, i, j, z)
        nh = 16 if d == 128 else (8 if d == 256 else 32)
        slot = torch.full((m,), -1, dtype=torch.int8, device=G.DEV)
        items = torch.arange(nh, dtype=torch.int32, device=G.DEV) + 2          # items 2 .. nh+1 are privatised
        slot[2:2 + nh] = torch.arange(nh, dtype=torch.int8, device=G.DEV)
        sl = slice(3, 3 + B)
        lo, gUo, gVo = O.loss_and_grads(U, V, u[sl], i[sl], j[sl], z[sl].astype(np.float32))
        Ud, Vd = G.dev_f32(U), G.dev_f32(V)
        for hot in (True, False):
            gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd); loss = torch.zeros(1, device=G.DEV)
            check(lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(store.rec), None, 3, B, d, 1.0 / B, ptr(gU),
                                              ptr(gV), ptr(loss), ptr(slot) if hot else None,
                                              ptr(items) if hot else None, nh if hot else 0, 1, current_stream()), "ex")
            assert abs(loss.item() - lo) < 2e-5 * abs(lo), (one_user, hot)
            assert G.rel(gU.cpu().numpy(), gUo) < 2e-5 and G.rel(gV.cpu().numpy(), gVo) < 2e-5, (one_user, hot)


@pytest.mark.parametrize("mode", ["deterministic", "deterministic_sort", "deterministic_fixed"])
def test_deterministic_mode_is_bit_reproducible(G, mode):
    rng = np.random.default_rng(7)
    U, V, u, i, j, z = _random_problem(rng, 500, 64, 64, 20000, hot=True)
    store = G.store_from(u, i, j, z)
    a = G.fwd_bwd(U, V, store, mode=mode)
    b = G.fwd_bwd(U, V, store, mode=mode)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    c = G.fwd_bwd(U, V, store, mode="atomic")
    assert G.rel(c[1], a[1]) < 1e-4 and G.rel(c[2], a[2]) < 1e-4 and abs(c[0] - a[0]) < 1e-5 * abs(a[0])
    if mode == "deterministic_fixed":
        # fixed-point sums do not depend on the order of the batch either: a permuted batch gives the same bits
        perm = rng.permutation(len(u))
        p = G.fwd_bwd(U, V, G.store_from(u[perm], i[perm], j[perm], z[perm]), mode=mode)
        assert a[0] == p[0] and np.array_equal(a[1], p[1]) and np.array_equal(a[2], p[2])
        # accumulation into non-empty gradient tables and a start offset
        lo, gUo, gVo = O.loss_and_grads(U, V, u[100:], i[100:], j[100:], z[100:].astype(np.float32))
        g0U = rng.standard_normal(U.shape).astype(np.float32); g0V = rng.standard_normal(V.shape).astype(np.float32)
        q = G.fwd_bwd(U, V, store, start=100, mode=mode, gU=g0U, gV=g0V)
        assert abs(q[0] - lo) < 1e-5 * abs(lo)
        assert np.abs(q[1] - (g0U + gUo)).max() < 2e-6 and np.abs(q[2] - (g0V + gVo)).max() < 2e-6


def test_gradients_accumulate_and_scale(G):
    """inv_batch is the global-batch scale (data parallel): two half batches with 1/B add up to the full batch."""
    rng = np.random.default_rng(9)
    U, V, u, i, j, z = _random_problem(rng, 50, 40, 16, 600)
    store = G.store_from(u, i, j, z)
    full = G.fwd_bwd(U, V, store, mode="deterministic")
    for mode in ("atomic", "deterministic"):
        l1, gU1, gV1 = G.fwd_bwd(U, V, store, start=0, B=300, mode=mode, inv_batch=1 / 600)
        l2, gU2, gV2 = G.fwd_bwd(U, V, store, start=300, B=300, mode=mode, inv_batch=1 / 600, gU=gU1, gV=gV1)
        assert abs(l1 + l2 - full[0]) < 1e-5 * abs(full[0])
        assert G.rel(gU2, full[1]) < 1e-5 and G.rel(gV2, full[2]) < 1e-5


def _adam_gpu(G, p, g, m, v, step, lr, wd, zero=1):
    from mfcd_b200._lib import lib, check, ptr, current_stream
    pd, gd, md, vd = G.dev_f32(p), G.dev_f32(g), G.dev_f32(m), G.dev_f32(v)
    check(lib.mfcd_adam_update(ptr(pd), ptr(gd), ptr(md), ptr(vd), pd.numel(), lr, 0.9, 0.999, 1e-8, wd, step, zero,
                               current_stream()), "adam")
    return pd.cpu().numpy(), gd.cpu().numpy(), md.cpu().numpy(), vd.cpu().numpy()


@pytest.mark.parametrize("numel", [1, 7, 4096, 100003])
def test_adam_kernel_matches_oracle(G, numel):
    rng = np.random.default_rng(numel)
    p = rng.standard_normal(numel).astype(np.float32)
    g = (rng.standard_normal(numel) * 1e-2).astype(np.float32)
    m = (rng.standard_normal(numel) * 1e-3).astype(np.float32)
    v = (rng.random(numel) * 1e-5).astype(np.float32)
    for step, wd in ((1, 0.0), (3, 1e-5), (1000, 1e-2)):
        po, mo, vo = O.adam_step(p.copy(), g.copy(), m.copy(), v.copy(), step, lr=1e-3, weight_decay=wd)
        pg, gg, mg, vg = _adam_gpu(G, p, g, m, v, step, 1e-3, wd)
        assert G.rel(pg, po) < 1e-6 and G.rel(mg, mo) < 1e-6 and G.rel(vg, vo) < 1e-5   # fma contraction on v
        assert not gg.any()                                   # zero_grad fused
    _, gkeep, _, _ = _adam_gpu(G, p, g, m, v, 1, 1e-3, 0.0, zero=0)
    assert np.array_equal(gkeep, g)


def test_adam_fixed_point(G):
    """zero gradient, zero state, no weight decay: parameters must not move (idempotence)."""
    p = np.linspace(-1, 1, 1001).astype(np.float32)
    z = np.zeros_like(p)
    pg, _, mg, vg = _adam_gpu(G, p, z, z, z, 1, 1e-3, 0.0)
    assert np.array_equal(pg, p) and not mg.any() and not vg.any()


def test_sgd_kernel_matches_oracle(G):
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(5)
    p = rng.standard_normal(1000).astype(np.float32)
    buf_o = np.zeros_like(p); pd = G.dev_f32(p); bd = G.dev_f32(buf_o); po = p.copy()
    for step in (1, 2, 3):
        g = rng.standard_normal(1000).astype(np.float32)
        O.sgd_step(po, g.copy(), buf_o, step, lr=0.05, momentum=0.9, weight_decay=1e-4)
        gd = G.dev_f32(g)
        check(lib.mfcd_sgd_update(ptr(pd), ptr(gd), ptr(bd), 1000, 0.05, 0.9, 1e-4, step, 1, current_stream()), "sgd")
        assert G.rel(pd.cpu().numpy(), po) < 1e-6 and not gd.cpu().numpy().any()


def _train_through_abi(G, g, mode, epochs=None):
    """Replay the reference's recorded batches through mfcd_train_epoch."""
    from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch, resolve_mode
    n, m, d = int(g["n"]), int(g["m"]), int(g["d"])
    model = MatrixFactorization(n, m, d)
    with torch.no_grad():
        model.U.copy_(torch.from_numpy(g["U0"])); model.V.copy_(torch.from_numpy(g["V0"]))
    fs = model.flat_state(G.DEV)
    spec = OptimizerSpec.adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    spe = int(g["steps_per_epoch"])
    epochs = int(g["epochs"]) if epochs is None else epochs
    sizes = g["batch_sizes"]
    losses = []
    off = 0
    for e in range(epochs):
        cnt = int(sizes[e * spe:(e + 1) * spe].sum())
        store = G.store_from(g["batch_u"][off:off + cnt], g["batch_i"][off:off + cnt], g["batch_j"][off:off + cnt],
                             g["batch_z"][off:off + cnt])
        off += cnt
        losses += run_epoch(fs, store, None, 64, spec, resolve_mode(mode, 64)).cpu().tolist()
    return model, np.array(losses)


@pytest.mark.parametrize("mode", ["deterministic", "atomic"])
@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_epoch_runner_matches_reference_steps(G, name, mode):
    """per-step loss and final U, V vs the reference's own run: 1e-5 relative (deterministic mode bar);
    the atomic mode is held to the same bar at these sizes."""
    g = load_golden(name)
    model, losses = _train_through_abi(G, g, mode)
    assert G.rel(losses, g["step_losses"]) < 1e-5
    assert G.rel(model.U.detach().cpu().numpy(), g["U1"]) < 1e-5
    assert G.rel(model.V.detach().cpu().numpy(), g["V1"]) < 1e-5


@pytest.mark.parametrize("n,m,d,steps,batch,lr,wd", [
    (1500, 700, 10, 150, 64, 1e-2, 1e-4), (1000, 1000, 2, 120, 64, 1e-2, 1e-4), (3000, 2000, 32, 60, 64, 1e-2, 1e-4),
    (400, 300, 64, 41, 200, 1e-2, 1e-4),
    # rows that are hardly ever touched only follow the weight decay; at lr = 1e-2 their Adam quotient
    # m / (sqrt(v) + eps) of tiny numbers drifts 6e-6 from numpy's in 80 steps on BOTH GPU paths (identical
    # values, MFCD_EPOCH_KERNEL=0/1), so this case runs at the reference's own lr / wd
    (5000, 4000, 8, 80, 7, 1e-3, 1e-5)])
def test_persistent_epoch_kernel_many_ctas_vs_oracle(G, n, m, d, steps, batch, lr, wd):
    """The one-kernel epoch (tables split over several CTAs, ping-pong parameters, grid barrier per step):
    per-step loss, U, V and a SECOND epoch (odd/even step counts, Adam state carried over, permuted order)
    against the numpy oracle, 1e-5 relative.  Ragged last batch included."""
    from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch
    rng = np.random.default_rng(n + d)
    N = steps * batch - 3
    U0 = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    V0 = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    u, i, j = rng.integers(0, n, N), rng.integers(0, m, N), rng.integers(0, m, N)
    u[:40] = 3; i[:20] = 5; j[20:40] = 5                       # repeated rows inside one batch, i and j sides
    z = rng.integers(0, 2, N).astype(np.float64)
    perm2 = rng.permutation(N)
    model = MatrixFactorization(n, m, d)
    with torch.no_grad():
        model.U.copy_(torch.from_numpy(U0)); model.V.copy_(torch.from_numpy(V0))
    fs = model.flat_state(G.DEV)
    spec = OptimizerSpec.adam(lr=lr, weight_decay=wd)
    store = G.store_from(u, i, j, z)
    l1 = run_epoch(fs, store, None, batch, spec, 1).cpu().numpy()
    l2 = run_epoch(fs, store, torch.from_numpy(perm2).to(G.DEV, torch.int32), batch, spec, 1).cpu().numpy()
    Uo, Vo = U0.copy(), V0.copy()
    lo1, st = O.train_steps(Uo, Vo, O.split_batches(u, i, j, z, batch), lr, wd)
    lo2, st = O.train_steps(Uo, Vo, O.split_batches(u, i, j, z, batch, order=perm2), lr, wd, state=st)
    assert G.rel(l1, lo1) < 1e-5 and G.rel(l2, lo2) < 1e-5
    assert G.rel(model.U.detach().cpu().numpy(), Uo) < 1e-5 and G.rel(model.V.detach().cpu().numpy(), Vo) < 1e-5
    assert not fs.grads.any().item()                            # the gradient buffer is never touched
    assert fs.step == 2 * steps


@pytest.mark.parametrize("name", TRAIN_FIXTURES)
def test_step_snapshots(G, name):
    from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch
    g = load_golden(name)
    n, m, d = int(g["n"]), int(g["m"]), int(g["d"])
    model = MatrixFactorization(n, m, d)
    with torch.no_grad():
        model.U.copy_(torch.from_numpy(g["U0"])); model.V.copy_(torch.from_numpy(g["V0"]))
    fs = model.flat_state(G.DEV)
    spec = OptimizerSpec.adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    batches = batches_from(g)
    done = 0
    for k in (1, 2, 5):
        for (u, i, j, z) in batches[done:k]:
            run_epoch(fs, G.store_from(u, i, j, z), None, 64, spec, 1)
        done = k
        assert fs.step == k
        assert G.rel(model.U.detach().cpu().numpy(), g[f"U_step{k}"]) < 1e-5
        assert G.rel(model.V.detach().cpu().numpy(), g[f"V_step{k}"]) < 1e-5


def test_config2_recorded_steps_through_the_persistent_epoch_kernel(G):
    """BASELINE config 2 (1000 x 1000, d = 10, K = 3, batch 64): the first 2048 steps the REFERENCE took
    (train_c2_steps.npz) replayed as one epoch of the persistent small-batch kernel (several CTAs, one grid barrier
    per step) -- per-step losses and weights within 1e-5; and the same through the two-launch path."""
    import ctypes as C
    from mfcd_b200 import _lib
    from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch
    g = load_golden("train_c2_steps.npz")
    n, m, d, steps = int(g["n"]), int(g["m"]), int(g["d"]), int(g["steps"])
    store = G.store_from(g["batch_u"].astype(np.int64), g["batch_i"].astype(np.int64), g["batch_j"].astype(np.int64),
                         g["batch_z"].astype(np.float64))
    spec = OptimizerSpec.adam(lr=float(g["lr"]), weight_decay=float(g["wd"]))
    for variant in ("persistent", "per-step launches"):
        model = MatrixFactorization(n, m, d)
        with torch.no_grad():
            model.U.copy_(torch.from_numpy(g["U0"])); model.V.copy_(torch.from_numpy(g["V0"]))
        fs = model.flat_state(G.DEV)
        if variant == "persistent":
            losses = run_epoch(fs, store, None, 64, spec, 1).cpu().numpy()
        else:                                       # no workspace -> k_det_small + k_adam per step (mfcd_b200.h)
            from mfcd_b200._lib import lib, check, ptr, current_stream
            out = torch.zeros(steps, dtype=torch.float32, device=G.DEV)
            a = _lib.EpochArgs()
            a.params, a.grads, a.state1, a.state2 = ptr(fs.params), ptr(fs.grads), ptr(fs.state1), ptr(fs.state2)
            a.n_users, a.n_items, a.d, a.optimizer, a.mode, a.flags = n, m, d, 0, 1, 0
            a.rec, a.perm, a.n_samples, a.batch_size = ptr(store.rec), None, len(store), 64
            a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, a.momentum = spec.lr, 0.9, 0.999, 1e-8, spec.weight_decay, 0.0
            a.step0, a.step_losses, a.workspace, a.workspace_bytes, a.stream = 0, ptr(out), None, 0, current_stream()
            check(lib.mfcd_train_epoch(C.byref(a)), "mfcd_train_epoch")
            fs.step += steps
            losses = out.cpu().numpy()
        assert len(losses) == steps and fs.step == steps
        assert np.abs(losses - g["step_losses"]).max() < 1e-5 * np.abs(g["step_losses"]).max(), variant
        assert G.rel(model.U.detach().cpu().numpy(), g["U_end"]) < 1e-5, variant
        assert G.rel(model.V.detach().cpu().numpy(), g["V_end"]) < 1e-5, variant


def test_train_model_api_with_recorded_order(G):
    """structure.train_model on loaders: epoch losses (mean of batch means) and validation losses vs the reference.
    The recorded epoch orders are injected through epoch_perm so the same batches are visited."""
    import structure
    from mfcd_b200.store import TripletLoader
    g = load_golden("train_d10_k3.npz")
    n, m, d = int(g["n"]), int(g["m"]), int(g["d"])
    train_store = G.store_from(g["train_u"], g["train_i"], g["train_j"], g["train_z"])
    val_store = G.store_from(g["val_u"], g["val_i"], g["val_j"], g["val_z"])
    # recover each epoch's permutation of the train dataset from the recorded batches
    key = lambda u, i, j, z: (u.astype(np.int64) * m + i) * m + j
    spe, N = int(g["steps_per_epoch"]), len(g["train_u"])
    perms = []
    ds_rows = np.stack([g["train_u"], g["train_i"], g["train_j"], g["train_z"]], 1)
    for e in range(int(g["epochs"])):
        rows = np.stack([g["batch_u"][e * N:(e + 1) * N], g["batch_i"][e * N:(e + 1) * N],
                         g["batch_j"][e * N:(e + 1) * N], g["batch_z"][e * N:(e + 1) * N]], 1)
        # rows may repeat (K=3 copies, equal labels): match greedily
        from collections import defaultdict
        pos = defaultdict(list)
        for idx, r in enumerate(map(tuple, ds_rows)):
            pos[r].append(idx)
        perms.append(np.array([pos[tuple(r)].pop() for r in rows], np.int64))

    class RecordedLoader(TripletLoader):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._e = 0

        def epoch_perm(self):
            p = torch.from_numpy(perms[self._e]).to(self.store.device, torch.int32)
            self._e += 1
            return p

    model = structure.MatrixFactorization(n, m, d)
    with torch.no_grad():
        model.U.copy_(torch.from_numpy(g["U0"])); model.V.copy_(torch.from_numpy(g["V0"]))
    opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=float(g["wd"]))
    tl = RecordedLoader(train_store, 64, shuffle=True)
    vl = TripletLoader(val_store, 64)
    tr, va = structure.train_model(model, tl, vl, opt, "cpu", num_epochs=int(g["epochs"]))
    assert G.rel(tr, g["train_losses"]) < 1e-5 and G.rel(va, g["val_losses"]) < 1e-5
    assert G.rel(model.U.detach().cpu().numpy(), g["U1"]) < 1e-5
    st = opt.state[model.U]
    assert int(st["step"]) == spe * int(g["epochs"]) and st["exp_avg"].shape == (n, d)
    # evaluation API
    test_store = G.store_from(g["test_u"], g["test_i"], g["test_j"], g["test_z"])
    loss, acc = structure.evaluate_model(model, TripletLoader(test_store, 64), "cpu")
    assert abs(loss - g["test_loss"]) < 1e-5 * abs(g["test_loss"]) and abs(acc - g["test_acc"]) < 1e-9
    gl, ga = structure.compute_ground_truth_metrics(TripletLoader(test_store, 64), torch.from_numpy(g["X"]), "cpu")
    assert abs(gl - g["gt_loss"]) < 1e-5 and abs(ga - g["gt_acc"]) < 1e-9
    with pytest.raises(ZeroDivisionError):
        structure.train_model(model, tl.__class__(train_store, 64), TripletLoader(val_store.slice(0, 0), 64), opt, "cpu",
                              num_epochs=1)


def test_second_train_model_call_with_a_fresh_optimizer_restarts_the_moments(G):
    """A fresh torch Adam starts from zero moments and step 0 (the reference builds one per repetition,
    structure.py:364); the same optimiser object passed again continues.  Also: device='cpu' hands CPU
    model.U / model.V back while the flat CUDA state stays current."""
    import structure
    from mfcd_b200.store import TripletLoader
    rng = np.random.default_rng(5)
    n, m, d, N = 40, 30, 8, 640
    u, i, j = rng.integers(0, n, N), rng.integers(0, m, N), rng.integers(0, m, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    store = G.store_from(u, i, j, z)
    val = G.store_from(u[:64], i[:64], j[:64], z[:64])
    torch.manual_seed(1)
    model = structure.MatrixFactorization(n, m, d)
    U0, V0 = model.U.detach().numpy().copy(), model.V.detach().numpy().copy()
    batches = O.split_batches(u, i, j, z, 64)
    lr, wd = 1e-2, 1e-4
    # oracle (updates its U / V arrays in place): epoch 1 with a fresh Adam, epoch 2 with ANOTHER fresh Adam,
    # epoch 3 continuing the second one
    Uo, Vo = U0.copy(), V0.copy()
    l1, _ = O.train_steps(Uo, Vo, batches, lr, wd)
    Ua = Uo.copy()
    l2, st2 = O.train_steps(Uo, Vo, batches, lr, wd)
    Ub, Vb = Uo.copy(), Vo.copy()
    l3, _ = O.train_steps(Uo, Vo, batches, lr, wd, state=st2)
    Uc = Uo.copy()
    tl, vl = TripletLoader(store, 64), TripletLoader(val, 64)
    opt1 = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    t1, _ = structure.train_model(model, tl, vl, opt1, "cpu", num_epochs=1)
    assert not model.U.is_cuda and not model.V.is_cuda                     # handed back on the caller's device
    assert G.rel(model.U.detach().numpy(), Ua) < 1e-5
    opt2 = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    t2, _ = structure.train_model(model, tl, vl, opt2, "cpu", num_epochs=1)
    assert abs(t2[0] - np.mean(l2)) < 1e-5 * abs(np.mean(l2))
    assert G.rel(model.U.detach().numpy(), Ub) < 1e-5 and G.rel(model.V.detach().numpy(), Vb) < 1e-5
    assert int(opt2.state[model.U]["step"]) == len(batches)
    t3, _ = structure.train_model(model, tl, vl, opt2, "cpu", num_epochs=1)  # same optimiser: continues
    assert abs(t3[0] - np.mean(l3)) < 1e-5 * abs(np.mean(l3))
    assert G.rel(model.U.detach().numpy(), Uc) < 1e-5
    assert int(opt2.state[model.U]["step"]) == 2 * len(batches)


def test_pack_unpack_gather_roundtrip(G):
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(1)
    N = 10007
    u, i, j = rng.integers(0, 2 ** 31 - 1, N), rng.integers(0, 1000, N), rng.integers(0, 1000, N)
    z = rng.integers(0, 5, N) / 4.0
    store = G.store_from(u, i, j, z)
    cu, ci, cj, cz = [c.cpu().numpy() for c in store.columns()]
    assert np.array_equal(cu, u) and np.array_equal(ci, i) and np.array_equal(cj, j) and np.array_equal(cz, z)
    perm = torch.randperm(N, device=G.DEV, dtype=torch.int32)
    out = torch.empty_like(store.rec)
    check(lib.mfcd_gather_triplets(ptr(store.rec), ptr(perm), N, ptr(out), current_stream()), "gather")
    assert torch.equal(out, store.rec[perm.long()])
    assert len(G.store_from([], [], [], [])) == 0
    # 8-byte wire format (hard labels): exact round trip at the index limits, loud failure otherwise
    from mfcd_b200._lib import MfcdError
    from mfcd_b200.store import TripletStore
    hard = G.store_from(rng.integers(0, 2 ** 23, N), rng.integers(0, 2 ** 20, N), rng.integers(0, 2 ** 20, N),
                        rng.integers(0, 2, N).astype(np.float64))
    assert torch.equal(TripletStore.from_packed8(hard.pack8()).rec, hard.rec)
    with pytest.raises(MfcdError):
        store.pack8()                      # soft labels and 31-bit user ids do not fit


# ---- properties at BASELINE.json's full table shape (config 4: 100k x 50k, d = 64) ------------------
def test_full_size_properties_c4_shape(G):
    from mfcd_b200._lib import lib, check, ptr, current_stream
    n, m, d, B = 100_000, 50_000, 64, 1 << 20
    gen = torch.Generator(device=G.DEV); gen.manual_seed(4)
    U = torch.randn(n, d, generator=gen, device=G.DEV) / 8
    V = torch.randn(m, d, generator=gen, device=G.DEV) / 8
    rec = torch.empty((B, 4), dtype=torch.int32, device=G.DEV)
    rec[:, 0] = torch.randint(0, n, (B,), generator=gen, device=G.DEV)
    rec[:, 1] = torch.randint(0, m, (B,), generator=gen, device=G.DEV)
    rec[:, 2] = torch.randint(0, m, (B,), generator=gen, device=G.DEV)
    rec[:, 3] = torch.randint(0, 2, (B,), generator=gen, device=G.DEV).float().view(torch.int32)
    outs = {}
    for mode in ("atomic", "deterministic"):
        gU = torch.zeros_like(U); gV = torch.zeros_like(V); loss = torch.zeros(1, device=G.DEV)
        if mode == "atomic":
            check(lib.mfcd_triplet_fwd_bwd(ptr(U), ptr(V), ptr(rec), None, 0, B, d, 1.0 / B, ptr(gU), ptr(gV), ptr(loss),
                                           current_stream()), "k1")
        else:
            need = C.c_size_t(0)
            check(lib.mfcd_det_workspace_bytes(B, d, C.byref(need)), "ws")
            ws = torch.empty(need.value, dtype=torch.uint8, device=G.DEV)
            check(lib.mfcd_triplet_fwd_bwd_det(ptr(U), ptr(V), ptr(rec), None, 0, B, d, 1.0 / B, n, m, ptr(gU), ptr(gV),
                                               ptr(loss), ptr(ws), need.value, current_stream()), "k1det")
        outs[mode] = (gU, gV, loss.item())
        # checksum of checksums: every triplet adds +g U_u to row i and -g U_u to row j => column sums of gV vanish
        col = gV.double().sum(dim=0).abs().max().item()
        assert col < 1e-6 * gV.double().abs().sum().item() / d + 1e-9
    # the two scatter modes agree at full size
    a, b = outs["atomic"], outs["deterministic"]
    assert (a[0] - b[0]).abs().max().item() < 1e-4 * b[0].abs().max().item()
    assert (a[1] - b[1]).abs().max().item() < 1e-4 * b[1].abs().max().item()
    assert abs(a[2] - b[2]) < 1e-5 * abs(b[2])
    # the loss equals the evaluation kernel's loss over the same records (one batch)
    bl = torch.zeros(1, device=G.DEV); correct = torch.zeros(2, dtype=torch.int64, device=G.DEV)
    check(lib.mfcd_triplet_eval(ptr(U), ptr(V), ptr(rec), B, d, B, ptr(bl), ptr(correct), ptr(correct[1:]),
                                current_stream()), "k4")
    assert abs(bl.item() - b[2]) < 1e-4 * abs(b[2])
    # linearity: gradients scale with inv_batch
    gU2 = torch.zeros_like(U); gV2 = torch.zeros_like(V); l2 = torch.zeros(1, device=G.DEV)
    check(lib.mfcd_triplet_fwd_bwd(ptr(U), ptr(V), ptr(rec), None, 0, B, d, 2.0 / B, ptr(gU2), ptr(gV2), ptr(l2),
                                   current_stream()), "k1")
    assert (gU2 - 2 * a[0]).abs().max().item() < 1e-4 * gU2.abs().max().item()
