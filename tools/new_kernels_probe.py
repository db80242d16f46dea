"""Small end-to-end exercise of the kernels written this round, (compute-sanitizer is closed on this GPU pool, so bounds are checked by small cases against the oracle):
K1 lean (hot rows + user runs + run-length wire input), group_by_user, pack/unpack wire, the persistent epoch
kernel on a 2-CTA cluster, and the atomic epoch path.  Prints 'probe ok' when the numbers also match the oracle."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import mfcd_b200
from mfcd_b200._lib import lib, check, ptr, current_stream
from mfcd_b200.store import TripletStore
from mfcd_b200.trainer import MatrixFactorization, OptimizerSpec, run_epoch
from oracle import mfcd_oracle as O

dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
n, m, d, B = 60, 90, 64, 3000
U = (rng.standard_normal((n, d)) / 8).astype(np.float32); V = (rng.standard_normal((m, d)) / 8).astype(np.float32)
pr = 1.0 / np.arange(1, m + 1) ** 1.5; pr /= pr.sum()
u = rng.integers(0, n, B); i = rng.choice(m, B, p=pr); j = rng.choice(m, B, p=pr)
keep = i != j; u, i, j = u[keep], i[keep], j[keep]; B = len(u)
z = rng.integers(0, 2, B).astype(np.float64)
lo, gUo, gVo = O.loss_and_grads(U, V, u, i, j, z.astype(np.float32))
store = TripletStore.from_columns(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(j), torch.from_numpy(z), device=dev)
store.group_by_user(B)
hot = store.hot_items(m, d, B, min_hits_per_batch=100)
wire = store.pack_wire(0, B)
assert torch.equal(TripletStore.from_wire(wire, B).rec, store.rec)
Ud, Vd = torch.from_numpy(U).to(dev), torch.from_numpy(V).to(dev)
for flags, rec in ((1, store.rec), (3, wire)):
    gU = torch.zeros_like(Ud); gV = torch.zeros_like(Vd); loss = torch.zeros(1, device=dev)
    check(lib.mfcd_triplet_fwd_bwd_ex(ptr(Ud), ptr(Vd), ptr(rec), None, 0, B, d, 1.0 / B, ptr(gU), ptr(gV), ptr(loss),
                                      ptr(hot[0]), ptr(hot[1]), hot[1].numel(), flags, current_stream()), "k1")
    assert abs(loss.item() - lo) < 2e-5 * abs(lo)
    assert np.abs(gU.cpu().numpy() - gUo).max() < 2e-5 * np.abs(gUo).max()
    assert np.abs(gV.cpu().numpy() - gVo).max() < 2e-5 * np.abs(gVo).max()
# persistent epoch kernel: 3200 elements -> 2 CTAs (one cluster), 20 steps of 64 + a ragged one
n2, m2, d2 = 200, 200, 8
N = 20 * 64 + 5
u2, i2, j2 = rng.integers(0, n2, N), rng.integers(0, m2, N), rng.integers(0, m2, N)
z2 = rng.integers(0, 2, N).astype(np.float64)
st2 = TripletStore.from_columns(torch.from_numpy(u2), torch.from_numpy(i2), torch.from_numpy(j2), torch.from_numpy(z2), device=dev)
for mode, batch in ((1, 64), (0, 512)):
    torch.manual_seed(0)
    model = MatrixFactorization(n2, m2, d2)
    U0, V0 = model.U.detach().numpy().copy(), model.V.detach().numpy().copy()
    fs = model.flat_state(dev)
    losses = run_epoch(fs, st2, None, batch, OptimizerSpec.adam(lr=1e-3, weight_decay=1e-5), mode).cpu().numpy()
    ref, _ = O.train_steps(U0, V0, O.split_batches(u2, i2, j2, z2, batch), 1e-3, 1e-5)
    assert np.abs(losses - np.array(ref)).max() < 1e-5 * np.abs(ref).max(), mode
    assert np.abs(model.U.detach().cpu().numpy() - U0).max() < 1e-5 * np.abs(U0).max() + 1e-6
torch.cuda.synchronize()
print("probe ok")
