"""CPU oracle for the triplet-comparison matrix-factorisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` are allowed to.  The product path is
the CUDA library behind ``include/mfcd_b200.h`` and it fails loudly when that
library is missing.

This is a plain-numpy restatement (fp32 arithmetic, explicit formulas, no
autograd) of what the reference computes through PyTorch ATen ops.  Every
function cites the reference lines it follows (paths relative to
``/root/reference``).  The arithmetic itself lives in third-party PyTorch
(unpinned in the reference's requirements.txt; 2.11.0 in this image), numpy and
scipy, so the formulas of ``binary_cross_entropy{,_backward}`` and
``torch.optim.Adam`` (single-tensor path) are restated from their published
definitions.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py``
(which imports ``/root/reference/structure.py``) and committed under
``tests/golden/*.npz``.  ``tests/test_oracle_vs_golden.py`` checks every
function below against those fixtures.
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# a2: forward  (structure.py:773-795)
# ----------------------------------------------------------------------------
def score_diff(U, V, u, i, j):
    """x_b = sum_k U[u_b,k] * (V[i_b,k] - V[j_b,k])   (structure.py:787-792).

    There is no scale ``s`` in the model's forward."""
    U = np.asarray(U, F32)
    V = np.asarray(V, F32)
    prod = U[u] * (V[i] - V[j])
    return prod.sum(axis=1, dtype=F32)


def sigmoid(x):
    """fp32 logistic, 1/(1+exp(-x))   (structure.py:795, torch.sigmoid)."""
    x = np.asarray(x, F32)
    with np.errstate(over="ignore"):
        return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def forward(U, V, u, i, j):
    return sigmoid(score_diff(U, V, u, i, j))


# ----------------------------------------------------------------------------
# a3: BCE loss + its gradient w.r.t. the score  (structure.py:849, :864, :908)
# ----------------------------------------------------------------------------
def bce_per_sample(p, z):
    """-[z*max(log p,-100) + (1-z)*max(log1p(-p),-100)]  (ATen binary_cross_entropy)."""
    p = np.asarray(p, F32)
    z = np.asarray(z, F32)
    with np.errstate(divide="ignore"):
        lp = np.maximum(np.log(p, dtype=F32), F32(-100))
        lq = np.maximum(np.log1p(-p, dtype=F32), F32(-100))
    return ((z - F32(1)) * lq - z * lp).astype(F32)


def bce_mean(p, z):
    """Batch-mean BCE as F.binary_cross_entropy(pred, z.float()) returns it."""
    per = bce_per_sample(p, z)
    return F32(per.astype(np.float64).mean())


def bce_grad_score(p, z, batch):
    """dL/dx_b for L = mean BCE(sigmoid(x)):  ATen binary_cross_entropy_backward
    followed by sigmoid_backward:

        g_in = (1/B) * (p - z) / max((1-p)*p, 1e-12)
        g_x  = g_in * (1-p) * p

    which equals (p-z)/B except that it is exactly 0 once p has saturated to
    0.0 or 1.0 in fp32."""
    p = np.asarray(p, F32)
    z = np.asarray(z, F32)
    q = (F32(1) - p) * p
    g_in = (F32(1) / F32(batch)) * (p - z) / np.maximum(q, F32(1e-12))
    return (g_in * (F32(1) - p) * p).astype(F32)


# ----------------------------------------------------------------------------
# a4: backward into dense gradients  (structure.py:850, autograd of :787-792)
# ----------------------------------------------------------------------------
def dense_grads(U, V, u, i, j, gx):
    """gU[u_b] += g_b (V[i_b]-V[j_b]); gV[i_b] += g_b U[u_b]; gV[j_b] -= g_b U[u_b].

    Accumulated in batch order into zero-initialised dense tables, which is the
    order the CPU ``index_put_(accumulate=True)`` of the reference uses."""
    U = np.asarray(U, F32)
    V = np.asarray(V, F32)
    gx = np.asarray(gx, F32)
    gU = np.zeros_like(U)
    gVi = np.zeros_like(V)
    gVj = np.zeros_like(V)
    du = gx[:, None] * (V[i] - V[j])
    dv = gx[:, None] * U[u]
    for b in range(len(gx)):  # sequential, batch order
        gU[u[b]] += du[b]
        gVi[i[b]] += dv[b]
        gVj[j[b]] -= dv[b]
    return gU, (gVi + gVj).astype(F32)


def loss_and_grads(U, V, u, i, j, z):
    """One forward+backward of structure.py:848-850 on a batch."""
    p = forward(U, V, u, i, j)
    loss = bce_mean(p, z)
    gx = bce_grad_score(p, z, len(z))
    gU, gV = dense_grads(U, V, u, i, j, gx)
    return loss, gU, gV


# ----------------------------------------------------------------------------
# a5: torch.optim.Adam, single-tensor path  (structure.py:364, :851)
# ----------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8,
              weight_decay=0.0):
    """In-place dense Adam with coupled L2.  ``step`` is the 1-based step count.

        g <- g + wd*p ; m <- m + (g-m)(1-b1) ; v <- b2*v + (1-b2) g*g
        p <- p - (lr/(1-b1^t)) * m / ( sqrt(v)/sqrt(1-b2^t) + eps )

    Bias corrections are python float64, everything else fp32."""
    if weight_decay != 0.0:
        g = (g + F32(weight_decay) * p).astype(F32)
    m += (g - m) * F32(1.0 - beta1)
    v *= F32(beta2)
    v += F32(1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    denom = np.sqrt(v, dtype=F32) / F32(bc2_sqrt) + F32(eps)
    p += F32(-step_size) * (m / denom)
    return p, m, v


def sgd_step(p, g, buf, step, lr, momentum=0.0, weight_decay=0.0):
    """torch.optim.SGD (dampening 0, no nesterov): g += wd*p; buf = mu*buf + g; p -= lr*buf."""
    if weight_decay != 0.0:
        g = (g + F32(weight_decay) * p).astype(F32)
    if momentum != 0.0:
        if step == 1:
            buf[...] = g
        else:
            buf *= F32(momentum)
            buf += g
        g = buf
    p += F32(-lr) * g
    return p, buf


# ----------------------------------------------------------------------------
# a6: train_model  (structure.py:812-878)
# ----------------------------------------------------------------------------
class AdamState:
    def __init__(self, U, V):
        self.mU = np.zeros_like(U)
        self.vU = np.zeros_like(U)
        self.mV = np.zeros_like(V)
        self.vV = np.zeros_like(V)
        self.step = 0


def train_steps(U, V, batches, lr, weight_decay, state=None, betas=(0.9, 0.999), eps=1e-8):
    """Run the reference's inner loop (structure.py:845-852) over an explicit
    list of batches ``(u, i, j, z)``; returns the per-step fp32 losses."""
    state = state or AdamState(U, V)
    losses = []
    for (u, i, j, z) in batches:
        loss, gU, gV = loss_and_grads(U, V, u, i, j, np.asarray(z, F32))
        state.step += 1
        adam_step(U, gU, state.mU, state.vU, state.step, lr, betas[0], betas[1], eps, weight_decay)
        adam_step(V, gV, state.mV, state.vV, state.step, lr, betas[0], betas[1], eps, weight_decay)
        losses.append(float(loss))
    return losses, state


def mean_of_batch_means(U, V, batches):
    """Validation / test loss as the reference reports it: the mean over
    batches of the batch-mean BCE (structure.py:858-868, :901-909, :921)."""
    tot = 0.0
    for (u, i, j, z) in batches:
        tot += float(bce_mean(forward(U, V, u, i, j), np.asarray(z, F32)))
    return tot / len(batches)


def split_batches(u, i, j, z, batch_size, order=None):
    n = len(u)
    order = np.arange(n) if order is None else np.asarray(order)
    out = []
    for s in range(0, n, batch_size):
        idx = order[s:s + batch_size]
        out.append((u[idx], i[idx], j[idx], z[idx]))
    return out


# ----------------------------------------------------------------------------
# a7: evaluate_model  (structure.py:881-921)
# ----------------------------------------------------------------------------
def evaluate_model(U, V, batches):
    loss_sum, correct, total = 0.0, 0, 0
    for (u, i, j, z) in batches:
        p = forward(U, V, u, i, j)
        loss_sum += float(bce_mean(p, np.asarray(z, F32)))
        hard = (p > F32(0.5)).astype(np.float64)
        correct += int((hard == np.asarray(z, np.float64)).sum())
        total += len(z)
    acc = correct / total if total > 0 else 0.0
    return loss_sum / len(batches), acc


# ----------------------------------------------------------------------------
# a8: compute_reconstruction_error  (structure.py:925-955)
# ----------------------------------------------------------------------------
def reconstruction_error(U, V, X, s):
    """|| (UV^T - column means) - sX ||_F / || sX ||_F  (column centring, sic)."""
    M = (np.asarray(U, F32) @ np.asarray(V, F32).T).astype(F32)
    M = M - M.mean(axis=0, keepdims=True, dtype=F32)
    sX = F32(s) * np.asarray(X, F32)
    num = math.sqrt(float(((M - sX).astype(np.float64) ** 2).sum()))
    den = math.sqrt(float((sX.astype(np.float64) ** 2).sum()))
    return num / den


# ----------------------------------------------------------------------------
# a9: compute_alpha_and_norm_ratios  (structure.py:958-1082)
# ----------------------------------------------------------------------------
def average_ranks(row):
    """1-based ranks with ties sharing their average rank (scipy rankdata 'average')."""
    row = np.asarray(row)
    order = np.argsort(row, kind="stable")
    srt = row[order]
    m = len(row)
    ranks = np.empty(m, np.float64)
    k = 0
    while k < m:
        e = k
        while e + 1 < m and srt[e + 1] == srt[k]:
            e += 1
        ranks[order[k:e + 1]] = 0.5 * (k + e) + 1.0
        k = e + 1
    return ranks


def pearson(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    a = a - a.mean()
    b = b - b.mean()
    den = math.sqrt(float((a * a).sum()) * float((b * b).sum()))
    return float((a * b).sum()) / den if den > 0 else float("nan")


def alpha_and_norm_ratios(U, V, X_init):
    """All 14 outputs of compute_alpha_and_norm_ratios, same order.

    Row centring of both UV^T and X (structure.py:985-987); alpha and norms
    (:990-996); per-row Pearson for rows whose std exceeds 1e-8 on both sides
    (:1003-1009); singular-value error (:1013-1017); per-row Spearman, NaNs
    dropped (:1024-1031); population std of the kept values (:1034-1035);
    slopes (:1039-1045); per-row alpha_i and the per-row-scaled error
    (:1054-1064)."""
    U = np.asarray(U, F32)
    V = np.asarray(V, F32)
    W = (U @ V.T).astype(F32)
    W = W - W.mean(axis=1, keepdims=True, dtype=F32)
    X = np.asarray(X_init, F32)
    X = X - X.mean(axis=1, keepdims=True, dtype=F32)
    W64, X64 = W.astype(np.float64), X.astype(np.float64)

    dot = float((W64 * X64).sum())
    norm_W = math.sqrt(float((W64 * W64).sum()))
    norm_X = math.sqrt(float((X64 * X64).sum()))
    alpha = dot / (norm_W ** 2 + 1e-8)
    norm_ratio = norm_W / (norm_X + 1e-8)
    rec_scaled = math.sqrt(float(((alpha * W64 - X64) ** 2).sum())) / (norm_X + 1e-8)

    n = X.shape[0]
    ok = [(np.std(X[r]) > 1e-8) and (np.std(W[r]) > 1e-8) for r in range(n)]
    correlations = [pearson(X[r], W[r]) for r in range(n) if ok[r]]
    pearson_mean = float(np.mean(correlations)) if correlations else 0.0

    S1 = np.linalg.svd(X64, compute_uv=False)
    S2 = np.linalg.svd(W64, compute_uv=False)
    k = min(len(S1), len(S2))
    svd_err = float(np.linalg.norm(alpha * S2[:k] - S1[:k]) / (np.linalg.norm(S1[:k]) + 1e-8))

    spearman_scores = []
    for r in range(n):
        if ok[r]:
            rho = pearson(average_ranks(X[r]), average_ranks(W[r]))
            if not math.isnan(rho):
                spearman_scores.append(rho)
    spearman_mean = float(np.mean(spearman_scores)) if spearman_scores else 0.0
    pearson_std = float(np.std(correlations)) if correlations else 0.0
    spearman_std = float(np.std(spearman_scores)) if spearman_scores else 0.0

    slopes = []
    for r in range(n):
        den = float(X64[r] @ X64[r])
        if den > 1e-8 and np.std(W[r]) > 1e-8:
            slopes.append(float(X64[r] @ W64[r]) / den)

    alpha_per_row = []
    adj = np.empty_like(W64)
    for r in range(n):
        den = float(W64[r] @ W64[r])
        a_r = float(X64[r] @ W64[r]) / den if den > 1e-8 else 0.0
        alpha_per_row.append(a_r)
        adj[r] = a_r * W64[r]
    rec_row = math.sqrt(float(((adj - X64) ** 2).sum())) / (norm_X + 1e-8)

    return (alpha, norm_X, norm_ratio, rec_scaled, pearson_mean, pearson_std,
            spearman_mean, spearman_std, svd_err, slopes, correlations,
            spearman_scores, rec_row, alpha_per_row)


# ----------------------------------------------------------------------------
# a10: compute_ground_truth_metrics  (structure.py:1085-1127)
# ----------------------------------------------------------------------------
def ground_truth_metrics(X, batches):
    """MSE(sigmoid(X[u,i]-X[u,j]), z) as mean of batch means (no scale s!) and
    accuracy of (diff>0)==z."""
    X = np.asarray(X, F32)
    loss_sum, correct, total = 0.0, 0, 0
    for (u, i, j, z) in batches:
        diff = X[u, i] - X[u, j]
        pr = sigmoid(diff)
        z32 = np.asarray(z, F32)
        loss_sum += float(F32(((pr - z32).astype(np.float64) ** 2).mean()))
        correct += int(((diff > 0).astype(np.float64) == np.asarray(z, np.float64)).sum())
        total += len(z)
    acc = correct / total if total > 0 else 0.0
    return loss_sum / len(batches), acc


# ----------------------------------------------------------------------------
# a11: BTLPreferenceDataset._generate_labels  (structure.py:493-519)
# ----------------------------------------------------------------------------
def btl_probability(X, u, i, j, scale):
    X = np.asarray(X, F32)
    return sigmoid(F32(scale) * (X[u, i] - X[u, j]))


def btl_labels_from_uniforms(q, uniforms, K, soft):
    """Labels from explicit uniforms in [0,1): draw k of triplet t is
    ``uniforms[t*K+k] < q[t]`` (torch.bernoulli == 'uniform < p').

    hard: K consecutive rows per triplet (structure.py:516-518);
    soft (train only): one row per triplet, label = mean of the K draws (:512)."""
    q = np.asarray(q, F32)
    draws = (np.asarray(uniforms, F32).reshape(len(q), K) < q[:, None]).astype(F32)
    if soft:
        return draws.mean(axis=1, dtype=F32)
    return draws.reshape(-1)


# ----------------------------------------------------------------------------
# a12: split_dataset_from_triplets sizes  (structure.py:703-730)
# ----------------------------------------------------------------------------
def split_sizes(total, train_ratio=0.8, val_ratio=0.1):
    tr = int(train_ratio * total)
    va = int(val_ratio * total)
    return tr, va, total - tr - va


def test_topup_needed(n_test, K, min_points=500):
    """How many extra triplets the test split receives (structure.py:721-724)."""
    if n_test * K < min_points:
        return (min_points + K - 1) // K - n_test
    return 0


# ----------------------------------------------------------------------------
# a13-a16: sampler acceptance rules  (generation_data.py:16-179)
# ----------------------------------------------------------------------------
def accept_stream(cands, num_triplets, exclude=(), predicate=None):
    """Sequential accept/reject shared by all four samplers: walk candidate
    triplets in order, keep (u,i,j) iff i != j, predicate holds, not excluded and
    not already kept; stop at ``num_triplets`` (generation_data.py:20-25, 69-78,
    122-127, 167-174).  Returns (kept list in acceptance order, candidates consumed)."""
    seen = set(exclude)
    kept = []
    used = 0
    for (u, i, j) in cands:
        if len(kept) >= num_triplets:
            break
        used += 1
        if i == j:
            continue
        if predicate is not None and not predicate(u, i, j):
            continue
        t = (int(u), int(i), int(j))
        if t in seen:
            continue
        seen.add(t)
        kept.append(t)
    return kept, used


def margin_threshold(X, num_triplets):
    """generation_data.py:56-57: mean over the first min(10,n) rows of
    (row max - row min), times num_triplets/(n*m)."""
    X = np.asarray(X, F32)
    n, m = X.shape
    sample = X[:min(10, n)]
    return float(np.mean(sample.max(axis=1) - sample.min(axis=1)) * num_triplets / (n * m))


def margin_predicate(X, margin):
    X = np.asarray(X, F32)
    return lambda u, i, j: abs(X[u, i] - X[u, j]) <= margin


def popularity_probs(m, method="zipf", alpha=1.5):
    """generation_data.py:110-119 (float64, over item index order)."""
    if method == "zipf":
        pr = 1.0 / (np.arange(1, m + 1) ** alpha)
    elif method == "exponential":
        pr = np.exp(-alpha * np.arange(m))
    elif method == "uniform":
        pr = np.ones(m)
    else:
        raise ValueError(f"Unknown popularity method: {method}")
    return pr / pr.sum()


def pair_without_replacement(probs, r1, r2):
    """Law of np.random.choice(m, 2, replace=False, p=probs) (generation_data.py:124)
    restated as two inverse-CDF draws: i ~ probs, then j ~ probs with item i
    removed and the rest renormalised.  (numpy's own algorithm consumes its
    uniforms differently, so this matches in distribution, not per-uniform.)"""
    cdf = np.cumsum(probs)
    i = int(min(np.searchsorted(cdf, r1 * cdf[-1], side="right"), len(probs) - 1))
    rest = probs.copy()
    rest[i] = 0.0
    cdf2 = np.cumsum(rest)
    j = int(min(np.searchsorted(cdf2, r2 * cdf2[-1], side="right"), len(probs) - 1))
    return i, j


def svd_rank(n, m, num_triplets):
    """generation_data.py:144: the rank actually used (argument is overridden)."""
    return int(num_triplets / (n * m) * max(n, m))


def svd_top_sets(X, rank, top_fraction=0.3):
    """generation_data.py:149-162: top users / items by row norm of U_k S_k and
    V_k S_k.  Uses a dense SVD truncated to ``rank`` (the reference uses ARPACK
    ``svds``; both give the same leading subspace up to round-off)."""
    X = np.asarray(X, np.float64)
    n, m = X.shape
    Uf, S, Vt = np.linalg.svd(X, full_matrices=False)
    Uk, Sk, Vk = Uf[:, :rank], S[:rank], Vt[:rank].T
    user_norms = np.linalg.norm(Uk * Sk, axis=1)
    item_norms = np.linalg.norm(Vk * Sk, axis=1)
    nu = max(1, int(top_fraction * n))
    ni = max(2, int(top_fraction * m))
    return np.argsort(user_norms)[-nu:], np.argsort(item_norms)[-ni:], user_norms, item_norms
