"""GPU tests of the product-API large-batch path: the per-epoch shuffle + user grouping (csrc/epoch_batches.cu),
train_model at throughput batch sizes against the oracle on the exact reshuffled batches, host-resident streaming
loaders in the three staging formats, and the fused exchange kernel with in-kernel flags on one GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import epoch_oracle as E
from oracle import mfcd_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("N", [1, 2, 3, 5, 64, 1000, 4097, 65536, 1_000_003])
def test_epoch_positions_are_a_permutation_and_match_the_oracle_bit_for_bit(G, N):
    from mfcd_b200._lib import lib, check, ptr, current_stream
    seed = 977 * N + 13
    pos = torch.empty(N, dtype=torch.int32, device=G.DEV)
    check(lib.mfcd_epoch_positions(N, seed, ptr(pos), current_stream()), "pos")
    got = pos.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, E.epoch_positions(N, seed))               # index work: bit-exact
    assert np.array_equal(np.sort(got), np.arange(N))                    # a permutation of [0, N)
    if N >= 1000:
        pos2 = torch.empty(N, dtype=torch.int32, device=G.DEV)
        check(lib.mfcd_epoch_positions(N, seed + 1, ptr(pos2), current_stream()), "pos")
        assert (pos2 != pos).float().mean().item() > 0.99               # another epoch, another order


def _records(rng, N, n, m):
    rec = np.empty((N, 4), np.int32)
    rec[:, 0] = np.sort(rng.integers(0, n, N))
    rec[:, 1], rec[:, 2] = rng.integers(0, m, N), rng.integers(0, m, N)
    rec[:, 3] = rng.integers(0, 2, N).astype(np.float32).view(np.int32)
    return rec


@pytest.mark.parametrize("N,B", [(1, 1), (31, 7), (4096, 512), (100_003, 4096), (70_000, 70_000), (300_000, 300),
                                 (2_000_000, 1 << 18), (500_000, 5000), (650_000, 10_000), (123_457, 1931)])
@pytest.mark.parametrize("given_pos", [False, True])
def test_epoch_batches_is_the_stable_multisplit_of_the_oracle(G, N, B, given_pos):
    """Every batch holds exactly the records whose epoch position falls in its range, in store order (so a
    user-sorted store gives user-grouped batches); both position sources; ragged last batch; up to 1000 batches."""
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(N + B)
    rec = _records(rng, N, 97, 50)
    d_rec = torch.from_numpy(rec).to(G.DEV)
    seed = 5 * N + B
    if given_pos:
        perm = rng.permutation(N)
        d_perm = torch.from_numpy(perm.astype(np.int32)).to(G.DEV)
        d_pos = torch.empty(N, dtype=torch.int32, device=G.DEV)
        check(lib.mfcd_invert_perm(ptr(d_perm), N, ptr(d_pos), current_stream()), "invert")
        pos = np.empty(N, np.int64); pos[perm] = np.arange(N)
        assert np.array_equal(d_pos.cpu().numpy(), pos)
    else:
        d_pos, pos = None, E.epoch_positions(N, seed)
    need = C.c_size_t(0)
    check(lib.mfcd_epoch_batches_workspace(N, B, C.byref(need)), "ws")
    ws = torch.empty(max(need.value, 1), dtype=torch.uint8, device=G.DEV)
    out = torch.full((N, 4), -1, dtype=torch.int32, device=G.DEV)
    check(lib.mfcd_epoch_batches(ptr(d_rec), N, B, ptr(d_pos), seed, ptr(out), ptr(ws), ws.numel(), current_stream()), "eb")
    want, batch = E.epoch_batches(rec, pos, B)
    got = out.cpu().numpy()
    assert np.array_equal(got, want)
    # inside every batch users are non-decreasing (the store was user-sorted), and sizes are exact
    for b in np.unique(batch)[:50]:
        rows = got[batch == b]
        assert (np.diff(rows[:, 0]) >= 0).all()
        assert len(rows) == min(B, N - b * B)


def test_epoch_batches_argument_errors(G):
    from mfcd_b200._lib import lib
    need = C.c_size_t(0)
    assert lib.mfcd_epoch_batches_workspace(10_000_000, 64, C.byref(need)) == -3      # too many batches: unsupported
    cap = C.c_int32(0)
    assert lib.mfcd_epoch_max_batches(C.byref(cap)) == 0 and cap.value >= 256


@pytest.mark.parametrize("rng_mode", ["device", "reference"])
def test_train_model_large_batch_equals_the_oracle_on_the_reshuffled_epochs(G, rng_mode):
    """structure.train_model at a throughput batch size (atomic scatter, per-epoch reshuffle + user grouping, hot
    rows): the oracle replays the exact batches of both epochs.  device RNG: batches rebuilt from the epoch seed;
    reference RNG: the reference's own RandomSampler permutation (same torch seed) cut into chunks of B."""
    import structure
    from mfcd_b200.store import TripletLoader
    rng = np.random.default_rng(21)
    n, m, d, B = 2000, 300, 16, 1 << 16
    N = 3 * B + 4097                                   # ragged last batch
    pr = 1.0 / np.arange(1, m + 1) ** 1.5; pr /= pr.sum()
    u = rng.integers(0, n, N); i = rng.choice(m, N, p=pr); j = (i + 1 + rng.choice(m - 1, N)) % m
    z = rng.integers(0, 2, N).astype(np.float64)
    store = G.store_from(u, i, j, z)
    val = G.store_from(u[:500], i[:500], j[:500], z[:500])
    torch.manual_seed(2)
    model = structure.MatrixFactorization(n, m, d)
    U0, V0 = model.U.detach().numpy().copy(), model.V.detach().numpy().copy()
    lr, wd = 5e-3, 1e-5
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    tl = TripletLoader(store, B, shuffle=True, shuffle_rng=rng_mode)
    vl = TripletLoader(val, B, shuffle=False, shuffle_rng=rng_mode)
    seeds = []
    orig = TripletLoader.epoch_records

    def spy(self, *a, **k):
        out = orig(self, *a, **k)
        seeds.append(self.last_epoch_seed)
        return out
    TripletLoader.epoch_records = spy
    try:
        torch.manual_seed(99)
        rng_state = torch.get_rng_state()
        t_losses, v_losses = structure.train_model(model, tl, vl, opt, "cuda", num_epochs=2, mode="atomic")
    finally:
        TripletLoader.epoch_records = orig
    assert len(seeds) == 2
    assert store.hot_items(m, d, B) is not None        # zipf items: the hot-row kernel ran
    Uo, Vo = U0.copy(), V0.copy()
    state, want = None, []
    if rng_mode == "device":
        rec = tl.store.rec.cpu().numpy()               # the loader re-ordered its store by user
        assert (np.diff(rec[:, 0]) >= 0).all()
    else:
        torch.set_rng_state(rng_state)
    for e in range(2):
        if rng_mode == "device":
            rows, batch = E.epoch_batches(rec, E.epoch_positions(N, seeds[e]), B)
        else:
            torch.empty((), dtype=torch.int64).random_()                     # DataLoader's _base_seed draw
            s = int(torch.empty((), dtype=torch.int64).random_().item())     # RandomSampler's seed
            perm = torch.randperm(N, generator=torch.Generator().manual_seed(s)).numpy()
            cols = np.stack([u, i, j], 1).astype(np.int32)
            rows = np.concatenate([cols[perm], z[perm].astype(np.float32).view(np.int32)[:, None]], 1)
            torch.empty((), dtype=torch.int64).random_()                     # the validation loader's iterator
        batches = [(r[:, 0].astype(np.int64), r[:, 1].astype(np.int64), r[:, 2].astype(np.int64),
                    r[:, 3].copy().view(np.float32).astype(np.float64)) for r in
                   (rows[s0:s0 + B] for s0 in range(0, N, B))]
        ls, state = O.train_steps(Uo, Vo, batches, lr, wd, state=state)
        want.append(float(np.mean(ls)))
    assert G.rel(t_losses, want) < 2e-5, (t_losses, want)
    assert G.rel(model.U.detach().cpu().numpy(), Uo) < 2e-5 and G.rel(model.V.detach().cpu().numpy(), Vo) < 2e-5
    vb = O.split_batches(u[:500], i[:500], j[:500], z[:500], B)
    assert abs(v_losses[-1] - O.mean_of_batch_means(Uo, Vo, vb)) < 1e-5


@pytest.mark.parametrize("fmt", ["records16", "wire8", "wire_rle", "wire8_live", "wire8_live:0.4", "wire8_live:auto"])
def test_host_resident_loader_streams_batches_and_matches_the_oracle(G, fmt):
    """train_model over a HostTripletLoader (pinned host batches, one H2D copy per step, loss read back per step):
    same steps as the oracle; the host packers are bit-identical to the device packers."""
    import structure
    from mfcd_b200 import hostpack
    from mfcd_b200.store import HostTripletLoader, TripletLoader, TripletStore
    rng = np.random.default_rng(8)
    n, m, d, B, N = 500, 120, 32, 5000, 5000 * 4 + 123
    u, i, j = rng.integers(0, n, N), rng.integers(0, m, N), rng.integers(0, m, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    rec = hostpack.as_records(u, i, j, z)
    fmt, _, frac = fmt.partition(":")            # wire8_live:f = only a fraction f of every batch is packed
    hl = HostTripletLoader.from_records(rec, B, fmt=fmt, pack_fraction=(frac or None) if frac in ("", "auto") else float(frac))
    assert len(hl) == 5 and hl.n_samples() == N and hl.sizes[-1] == 123
    # host packers == device packers, bit for bit
    dstore = G.store_from(u[:B], i[:B], j[:B], z[:B])
    assert np.array_equal(dstore.pack8().cpu().numpy().view(np.uint64), hostpack.pack8(rec[:B]))
    grouped = TripletStore(dstore.rec.clone()).group_by_user(B)
    assert np.array_equal(grouped.rec.cpu().numpy(), hostpack.group_by_user(rec[:B]))
    assert np.array_equal(grouped.pack_wire(0, B).cpu().numpy().view(np.uint32), hostpack.pack_wire(hostpack.group_by_user(rec[:B])))
    torch.manual_seed(6)
    model = structure.MatrixFactorization(n, m, d)
    U0, V0 = model.U.detach().numpy().copy(), model.V.detach().numpy().copy()
    lr, wd = 1e-2, 1e-5
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
    vl = TripletLoader(G.store_from(u[:300], i[:300], j[:300], z[:300]), 64)
    t_losses, _ = structure.train_model(model, hl, vl, opt, "cuda", num_epochs=2, mode="atomic")
    Uo, Vo = U0.copy(), V0.copy()
    batches = O.split_batches(u, i, j, z, B)
    l1, st = O.train_steps(Uo, Vo, batches, lr, wd)
    l2, st = O.train_steps(Uo, Vo, batches, lr, wd, state=st)
    assert G.rel(t_losses, [np.mean(l1), np.mean(l2)]) < 2e-5
    assert G.rel(model.U.detach().cpu().numpy(), Uo) < 2e-5 and G.rel(model.V.detach().cpu().numpy(), Vo) < 2e-5
    if fmt == "records16":      # deterministic scatter over streamed batches, too
        torch.manual_seed(6)
        model2 = structure.MatrixFactorization(n, m, d)
        opt2 = torch.optim.Adam(model2.parameters(), lr=lr, weight_decay=wd)
        t2, _ = structure.train_model(model2, hl, vl, opt2, "cuda", num_epochs=1, mode="deterministic")
        assert abs(t2[0] - np.mean(l1)) < 1e-5 * abs(np.mean(l1))
    with pytest.raises(ValueError):
        HostTripletLoader([torch.zeros((4, 4), dtype=torch.int32)], [4])       # not pinned


@pytest.mark.parametrize("numel", [4096, 1_000_003])
def test_fused_exchange_with_in_kernel_flags_on_one_rank(G, numel):
    """mfcd_dp_fused_adam_sync with world = 1 (the rank exchanges with itself through the same flag protocol):
    equals K3 on the same inputs, leaves the gradient buffer cleared, survives repeated calls (sequence numbers)."""
    from mfcd_b200._lib import lib, check, ptr, current_stream
    rng = np.random.default_rng(numel)
    p0 = rng.standard_normal(numel).astype(np.float32)
    p = G.dev_f32(p0); g = torch.zeros_like(p); m1 = torch.zeros_like(p); v1 = torch.zeros_like(p)
    pr = G.dev_f32(p0); gr = torch.zeros_like(p); mr = torch.zeros_like(p); vr = torch.zeros_like(p)
    flags = torch.zeros(64, dtype=torch.int32, device=G.DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=G.DEV)
    err = torch.zeros(1, dtype=torch.int32, device=G.DEV)
    arr = lambda t: (C.c_uint64 * 1)(t.data_ptr())
    for step in range(1, 4):
        gnp = rng.standard_normal(numel).astype(np.float32)
        g.copy_(torch.from_numpy(gnp)); gr.copy_(torch.from_numpy(gnp))
        check(lib.mfcd_dp_fused_adam_sync(arr(g), arr(p), arr(flags), 0, 0, 0, 1, numel, ptr(m1), ptr(v1), 1e-2, 0.9,
                                          0.999, 1e-8, 1e-4, step, step, ptr(counter), ptr(err), None, current_stream()), "k9")
        check(lib.mfcd_adam_update(ptr(pr), ptr(gr), ptr(mr), ptr(vr), numel, 1e-2, 0.9, 0.999, 1e-8, 1e-4, step, 1,
                                   current_stream()), "k3")
        torch.cuda.synchronize()
        assert int(err.item()) == 0 and int(counter.item()) == 0
        assert torch.equal(p, pr) and torch.equal(m1, mr) and torch.equal(v1, vr)
        assert not g.any()
    assert flags[0].item() == 3 and flags[8].item() == 3
    # double-buffered gradients: the call consumes one buffer (left as is) and clears the OTHER one locally
    padded = (numel + 3) // 4 * 4
    bufs = [torch.zeros(padded, device=G.DEV), torch.zeros(padded, device=G.DEV)]
    for step in range(4, 8):
        cur, other = bufs[step & 1], bufs[(step + 1) & 1]
        gnp = rng.standard_normal(numel).astype(np.float32)
        assert not cur.any()                                   # cleared by the previous call (or never used)
        cur[:numel].copy_(torch.from_numpy(gnp)); gr.copy_(torch.from_numpy(gnp))
        other.fill_(7.0)                                       # whatever the step before the previous one left
        check(lib.mfcd_dp_fused_adam_sync(arr(cur), arr(p), arr(flags), 0, 0, 0, 1, numel, ptr(m1), ptr(v1), 1e-2, 0.9,
                                          0.999, 1e-8, 1e-4, step, step, ptr(counter), ptr(err), ptr(other),
                                          current_stream()), "k9")
        check(lib.mfcd_adam_update(ptr(pr), ptr(gr), ptr(mr), ptr(vr), numel, 1e-2, 0.9, 0.999, 1e-8, 1e-4, step, 1,
                                   current_stream()), "k3")
        torch.cuda.synchronize()
        assert int(err.item()) == 0
        assert torch.equal(p, pr) and torch.equal(m1, mr) and torch.equal(v1, vr)
        assert not other.any() and torch.equal(cur[:numel].cpu(), torch.from_numpy(gnp))
        cur.zero_()
