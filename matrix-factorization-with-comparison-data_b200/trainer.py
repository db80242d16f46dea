"""Model, optimiser glue and the training / evaluation drivers of the hot path.

Mirrors (same names, arguments and return values) the reference's
``MatrixFactorization`` (structure.py:746-795), ``train_model`` (:812-878),
``evaluate_model`` (:881-921) and ``compute_ground_truth_metrics`` (:1085-1127);
the arithmetic runs in the CUDA kernels behind include/mfcd_b200.h.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, current_stream
from .store import GroundTruth, TripletLoader, TripletStore, as_loader, compute_device

MODE_ATOMIC = 0
MODE_DETERMINISTIC = 1


def resolve_mode(mode, batch_size):
    """'atomic' | 'deterministic' | 'auto' (env MFCD_MODE overrides 'auto').
    auto: deterministic for the reference's small batches (bit-reproducible like
    the CPU reference), atomic red.global scatter for throughput-sized batches."""
    if mode in (None, "auto"):
        mode = os.environ.get("MFCD_MODE", "auto")
    if mode == "auto":
        return MODE_DETERMINISTIC if batch_size <= 256 else MODE_ATOMIC
    if mode in ("atomic", MODE_ATOMIC):
        return MODE_ATOMIC
    if mode in ("deterministic", "det", MODE_DETERMINISTIC):
        return MODE_DETERMINISTIC
    raise ValueError(f"unknown scatter mode: {mode!r}")


class MatrixFactorization(nn.Module):
    """sigma(<U_u, V_i - V_j>) with U in R^{n x d}, V in R^{m x d}  (structure.py:746-795).

    Parameters are drawn exactly like the reference does (``randn(n,d)/sqrt(d)``
    from the global CPU generator, U first), so a seeded run starts from the same
    weights.  For training the two tables are packed into one flat CUDA buffer
    (U then V) of which ``self.U`` / ``self.V`` are views: one optimiser launch
    and one gradient all-reduce cover both.
    """

    def __init__(self, n_users, n_items, d):
        super().__init__()
        scale = torch.sqrt(torch.tensor(d, dtype=torch.float32))
        self.U = nn.Parameter(torch.randn(n_users, d) / scale)
        self.V = nn.Parameter(torch.randn(n_items, d) / scale)
        self._flat = None

    # -- flat CUDA storage ---------------------------------------------------
    def flat_state(self, device=None, storage=None):
        """storage = (params, grads): externally allocated flat fp32 CUDA buffers to live in."""
        n, d = self.U.shape
        m = self.V.shape[0]
        fs = self._flat
        if (storage is None and fs is not None and fs.params.is_cuda and self.U.data_ptr() == fs.params.data_ptr()
                and self.V.data_ptr() == fs.params.data_ptr() + 4 * n * d):
            return fs
        if storage is None and fs is not None and fs.host_mirror is not None and fs.host_mirror == self._mirror_key():
            return fs           # U / V are the untouched host copies mirror_to_host() left: the flat state is current
        dev = compute_device(device if device is not None else (self.U.device if self.U.is_cuda else None))
        fs = _FlatState(n, m, d, dev, *(storage or ()))
        with torch.no_grad():
            fs.params[: n * d].view(n, d).copy_(self.U.detach())
            fs.params[n * d:].view(m, d).copy_(self.V.detach())
        self.U.data = fs.params[: n * d].view(n, d)
        self.V.data = fs.params[n * d:].view(m, d)
        self._flat = fs
        return fs

    def _mirror_key(self):
        return (self.U.data_ptr(), self.U._version, self.V.data_ptr(), self.V._version)

    def mirror_to_host(self):
        """``device='cpu'`` callers get CPU ``model.U`` / ``model.V`` back, like the reference's
        ``model.to('cpu')`` would leave them (reference-style post-processing against a CPU X keeps working);
        the flat CUDA state stays attached and is reused as long as the host copies are not modified."""
        fs = self._flat
        if fs is None or not fs.params.is_cuda:
            return
        n, d, m = fs.n, fs.d, fs.m
        with torch.no_grad():
            self.U.data = fs.params[: n * d].view(n, d).cpu()
            self.V.data = fs.params[n * d:].view(m, d).cpu()
        fs.host_mirror = self._mirror_key()

    def forward(self, u, i, j):
        """Preference probabilities for index tensors u, i, j (inference only: the
        training path is the fused forward/backward kernel, not autograd)."""
        fs = self.flat_state()
        dev = fs.params.device
        u = torch.as_tensor(u).to(dev, torch.int64).contiguous()
        i = torch.as_tensor(i).to(dev, torch.int64).contiguous()
        j = torch.as_tensor(j).to(dev, torch.int64).contiguous()
        out = torch.empty(u.numel(), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.mfcd_triplet_scores(ptr(fs.U), ptr(fs.V), ptr(u), ptr(i), ptr(j), u.numel(), fs.d, ptr(out),
                                          current_stream()), "mfcd_triplet_scores")
        return out


class _FlatState:
    """(n+m)*d fp32 parameters, gradients and two optimiser moments, U first."""

    def __init__(self, n, m, d, device, params=None, grads=None):
        self.n, self.m, self.d = n, m, d
        numel = (n + m) * d
        # params / grads may be handed in (symmetric-memory buffers shared with peer GPUs, dist.PeerExchange)
        self.params = torch.zeros(numel, dtype=torch.float32, device=device) if params is None else params[:numel]
        self.grads = torch.zeros(numel, dtype=torch.float32, device=device) if grads is None else grads[:numel]
        self.state1 = torch.zeros(numel, dtype=torch.float32, device=device)
        self.state2 = torch.zeros(numel, dtype=torch.float32, device=device)
        self.step = 0
        self.workspace = None
        self.host_mirror = None

    @property
    def U(self):
        return self.params[: self.n * self.d]

    @property
    def V(self):
        return self.params[self.n * self.d:]

    def ensure_workspace(self, nbytes):
        if nbytes and (self.workspace is None or self.workspace.numel() < nbytes):
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.params.device)
        return self.workspace


class OptimizerSpec:
    """Hyper-parameters read off a torch.optim.Adam / SGD instance (the object
    run_experiment builds at structure.py:364)."""

    def __init__(self, optimizer):
        if isinstance(optimizer, OptimizerSpec):
            self.__dict__.update(optimizer.__dict__)
            return
        if len(optimizer.param_groups) != 1:
            raise NotImplementedError("one param group (the reference's setup) is supported")
        g = optimizer.param_groups[0]
        name = type(optimizer).__name__
        self.momentum = 0.0
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-8
        if name == "Adam":
            if g.get("amsgrad", False) or g.get("maximize", False) or g.get("decoupled_weight_decay", False):
                raise NotImplementedError("amsgrad / maximize / decoupled weight decay are not implemented")
            self.kind = 0
            self.beta1, self.beta2 = g["betas"]
            self.eps = g["eps"]
        elif name == "SGD":
            if g.get("nesterov", False) or g.get("dampening", 0) != 0 or g.get("maximize", False):
                raise NotImplementedError("nesterov / dampening / maximize are not implemented")
            self.kind = 1
            self.momentum = g.get("momentum", 0.0)
        else:
            raise NotImplementedError(f"optimizer {name} has no fused kernel (Adam and SGD do)")
        self.lr = float(g["lr"])
        self.weight_decay = float(g.get("weight_decay", 0.0))

    @classmethod
    def adam(cls, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        s = cls.__new__(cls)
        s.kind, s.lr, s.weight_decay, s.momentum = 0, float(lr), float(weight_decay), 0.0
        s.beta1, s.beta2, s.eps = betas[0], betas[1], eps
        return s


def _import_optimizer_state(model, optimizer, fs):
    """Continue from a torch optimiser that already holds state for U / V."""
    if isinstance(optimizer, OptimizerSpec) or not hasattr(optimizer, "state"):
        return
    n, d, m = fs.n, fs.d, fs.m
    if not optimizer.state.get(model.U, None) and not optimizer.state.get(model.V, None):
        # a fresh torch optimiser starts from zero moments and step 0 (the reference builds a new Adam per
        # repetition, structure.py:364) -- do not carry over what an earlier optimiser left in the flat state
        fs.state1.zero_()
        fs.state2.zero_()
        fs.step = 0
        return
    for p, sl in ((model.U, slice(0, n * d)), (model.V, slice(n * d, (n + m) * d))):
        st = optimizer.state.get(p, None)
        if not st:
            continue
        if "exp_avg" in st and st["exp_avg"].data_ptr() != fs.state1[sl].data_ptr():
            fs.state1[sl].copy_(st["exp_avg"].reshape(-1))
            fs.state2[sl].copy_(st["exp_avg_sq"].reshape(-1))
        if "momentum_buffer" in st and st["momentum_buffer"] is not None \
                and st["momentum_buffer"].data_ptr() != fs.state1[sl].data_ptr():
            fs.state1[sl].copy_(st["momentum_buffer"].reshape(-1))
        if "step" in st:
            fs.step = max(fs.step, int(float(st["step"])))


def _export_optimizer_state(model, optimizer, fs, spec):
    """Leave the torch optimiser object consistent with what the kernels did."""
    if isinstance(optimizer, OptimizerSpec) or not hasattr(optimizer, "state"):
        return
    n, d, m = fs.n, fs.d, fs.m
    for p, sl, shape in ((model.U, slice(0, n * d), (n, d)), (model.V, slice(n * d, (n + m) * d), (m, d))):
        if spec.kind == 0:
            optimizer.state[p] = {"step": torch.tensor(float(fs.step)),
                                  "exp_avg": fs.state1[sl].view(shape),
                                  "exp_avg_sq": fs.state2[sl].view(shape)}
        elif spec.momentum != 0.0:
            optimizer.state[p] = {"momentum_buffer": fs.state1[sl].view(shape)}


def run_epoch(fs: _FlatState, store: TripletStore, perm, batch_size, spec: OptimizerSpec, mode):
    """One epoch of structure.py:845-852 through mfcd_train_epoch; returns the
    per-step batch-mean losses as a float32 device tensor."""
    N = len(store)
    n_steps = (N + batch_size - 1) // batch_size
    dev = fs.params.device
    losses = torch.zeros(max(n_steps, 1), dtype=torch.float32, device=dev)
    if n_steps == 0:
        return losses[:0]
    a = _lib.EpochArgs()
    a.params, a.grads, a.state1, a.state2 = ptr(fs.params), ptr(fs.grads), ptr(fs.state1), ptr(fs.state2)
    a.n_users, a.n_items, a.d = fs.n, fs.m, fs.d
    a.optimizer, a.mode = spec.kind, mode
    a.flags = store.k1_flags(batch_size, perm) if mode == MODE_ATOMIC else 0
    a.rec, a.perm = ptr(store.rec), ptr(perm)
    a.n_samples, a.batch_size = N, batch_size
    a.lr, a.beta1, a.beta2, a.eps = spec.lr, spec.beta1, spec.beta2, spec.eps
    a.weight_decay, a.momentum = spec.weight_decay, spec.momentum
    a.step0 = fs.step
    a.step_losses = ptr(losses)
    ws_bytes = C.c_size_t(0)
    check(lib.mfcd_train_epoch_workspace(C.byref(a), C.byref(ws_bytes)), "mfcd_train_epoch_workspace")
    ws = fs.ensure_workspace(ws_bytes.value)
    a.workspace, a.workspace_bytes = ptr(ws), (ws.numel() if ws is not None else 0)
    hot = store.hot_items(fs.m, fs.d, batch_size) if mode == MODE_ATOMIC else None
    a.item_slot, a.hot_items, a.n_hot = (ptr(hot[0]), ptr(hot[1]), hot[1].numel()) if hot else (None, None, 0)
    with torch.cuda.device(dev):
        a.stream = current_stream()
        check(lib.mfcd_train_epoch(C.byref(a)), "mfcd_train_epoch")
    fs.step += n_steps
    return losses[:n_steps]


def eval_batches(fs: _FlatState, store: TripletStore, batch_size):
    """-> (per-batch mean BCE as float32 device tensor, #correct as int)."""
    N = len(store)
    nb = (N + batch_size - 1) // batch_size
    dev = fs.params.device
    batch_loss = torch.zeros(max(nb, 1), dtype=torch.float32, device=dev)
    correct = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.mfcd_triplet_eval(ptr(fs.U), ptr(fs.V), ptr(store.rec), N, fs.d, batch_size, ptr(batch_loss),
                                    ptr(correct), current_stream()), "mfcd_triplet_eval")
    return batch_loss[:nb], correct


def _sum_like_python(values):
    """`total += loss.item()` over float32 values, in double, in order."""
    tot = 0.0
    for v in values.tolist():
        tot += v
    return tot


def train_model(model, train_loader, val_loader, optimizer, device, num_epochs=100, is_last=False,
                open_browser=False, *, mode="auto", progress=False):
    """Same contract as the reference's train_model (structure.py:812-878):
    returns ``(train_losses, val_losses)``, one mean-of-batch-means per epoch.

    Per epoch the whole batch loop runs inside one C call (K1 + K3 per step, no
    per-step host sync; the reference syncs on ``loss.item()`` every step)."""
    dev = compute_device(device)
    train_loader = as_loader(train_loader, dev)
    val_loader = as_loader(val_loader, dev)
    fs = model.flat_state(dev)
    spec = OptimizerSpec(optimizer)
    _import_optimizer_state(model, optimizer, fs)
    scatter = resolve_mode(mode, train_loader.batch_size)

    train_losses, val_losses = [], []
    epochs = range(num_epochs)
    if progress:
        from tqdm import tqdm
        epochs = tqdm(epochs, desc="Training Progress")
    for _ in epochs:
        perm = train_loader.epoch_perm()
        step_losses = run_epoch(fs, train_loader.store, perm, train_loader.batch_size, spec, scatter)
        val_loader.begin_iteration()
        vloss, _ = eval_batches(fs, val_loader.store, val_loader.batch_size)
        # one host sync per epoch
        train_losses.append(_sum_like_python(step_losses) / len(train_loader))
        val_losses.append(_sum_like_python(vloss) / len(val_loader))      # ZeroDivisionError like the reference
    _export_optimizer_state(model, optimizer, fs, spec)
    if device is not None and torch.device(device).type == "cpu":
        model.mirror_to_host()
    return train_losses, val_losses


def evaluate_model(model, test_loader, device):
    """(mean-of-batch-means BCE, accuracy) on the test loader (structure.py:881-921)."""
    dev = compute_device(device)
    loader = as_loader(test_loader, dev)
    fs = model.flat_state(dev)
    loader.begin_iteration()
    batch_loss, correct = eval_batches(fs, loader.store, loader.batch_size)
    total = len(loader.store)
    accuracy = int(correct.item()) / total if total > 0 else 0.0
    return _sum_like_python(batch_loss) / len(loader), accuracy


def compute_ground_truth_metrics(test_loader, X, device):
    """(mean-of-batch-means MSE of sigmoid(X[u,i]-X[u,j]) vs label, accuracy of
    (diff > 0) == label) -- structure.py:1085-1127; note: no scale s."""
    dev = compute_device(device)
    loader = as_loader(test_loader, dev)
    gt = GroundTruth.wrap(X, dev)
    loader.begin_iteration()
    N = len(loader.store)
    nb = len(loader)
    batch_mse = torch.zeros(max(nb, 1), dtype=torch.float32, device=dev)
    correct = torch.zeros(1, dtype=torch.int64, device=dev)
    xv = gt.xview()
    with torch.cuda.device(dev):
        check(lib.mfcd_ground_truth_eval(C.byref(xv), ptr(loader.store.rec), N, loader.batch_size, ptr(batch_mse),
                                         ptr(correct), current_stream()), "mfcd_ground_truth_eval")
    accuracy = int(correct.item()) / N if N > 0 else 0.0
    return _sum_like_python(batch_mse[:nb]) / nb, accuracy
