"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the epoch batching of csrc/epoch_batches.cu.

The reference reshuffles with torch's RandomSampler (structure.py:738); in device RNG mode the product path
replaces the materialised permutation by a keyed bijection of [0, N) (a 3-round mixed-radix network over a domain
c * 2^k that hugs N, with cycle walking for the few values that leave [0, N)) and forms batch b = {r : pos(r) // B == b}, each batch in store order.
This file restates both so that the tests can check the kernels bit for bit (index work), and so that a test can
rebuild the exact batches an epoch visits and replay them through the training oracle.
Imported by tests/ only.
"""
import numpy as np

U32 = np.uint64(0xFFFFFFFF)


def round_keys(seed, count=12):
    """splitmix64 stream -> key words (bits 16..47 of each output)"""
    keys = []
    s = int(seed) & 0xFFFFFFFFFFFFFFFF
    for _ in range(count):
        s = (s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
        keys.append((z >> 16) & 0xFFFFFFFF)
    return keys


def domain(N):
    """(k, c): the bijection lives on [0, c * 2^k), the smallest such range >= N with c <= 64"""
    bits = 2
    while (1 << bits) < N:
        bits += 1
    k = max(bits // 2, bits - 6)
    return k, max(1, (N + (1 << k) - 1) >> k)


def _once(x, w, k, c):
    """one pass of the permutation of [0, c * 2^k): three rounds of (lo-mix keyed by hi, hi-add keyed by lo)"""
    mask = np.uint64((1 << k) - 1)
    sh = np.uint64(max(k >> 1, 1))
    kk, cc = np.uint64(k), np.uint64(c)
    hi, lo = x >> kk, x & mask
    for r in range(3):
        f = (((hi + np.uint64(1)) * np.uint64(w[3 * r] | 1)) & U32) >> np.uint64(32 - k)
        lo = ((lo ^ f) * np.uint64(((w[3 * r + 1] << 1) | 1) & 0xFFFFFFFF)) & mask
        lo = lo ^ (lo >> sh)
        lo = (lo + np.uint64(w[3 * r + 2])) & mask
        h = ((lo + np.uint64(w[9 + r])) * np.uint64(0x9E3779B1)) & U32
        h = h ^ (h >> np.uint64(15))
        h = (h * np.uint64(0x85EBCA77)) & U32
        hi = hi + ((h * cc) >> np.uint64(32))
        hi = np.where(hi >= cc, hi - cc, hi)
    return (hi << kk) | lo


def epoch_positions(N, seed):
    """pos[r] for r in [0, N): a permutation of [0, N)."""
    N = int(N)
    k, c = domain(N)
    w = round_keys(seed)
    x = _once(np.arange(N, dtype=np.uint64), w, k, c)
    todo = np.nonzero(x >= N)[0]
    while todo.size:
        x[todo] = _once(x[todo], w, k, c)
        todo = todo[x[todo] >= N]
    return x.astype(np.int64)


def epoch_batches(rec, pos, B):
    """stable multisplit: batch b = records with pos // B == b, in store order; batches concatenated"""
    batch = np.asarray(pos, np.int64) // int(B)
    order = np.argsort(batch, kind="stable")
    return rec[order], batch[order]
