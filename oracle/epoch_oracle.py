"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the epoch batching of csrc/epoch_batches.cu.

The reference reshuffles with torch's RandomSampler (structure.py:738); in device RNG mode the product path
replaces the materialised permutation by a keyed bijection of [0, N) (8-round alternating Feistel network over
ceil(log2 N) bits with cycle walking) and forms batch b = {r : pos(r) // B == b}, each batch in store order.
This file restates both so that the tests can check the kernels bit for bit (index work), and so that a test can
rebuild the exact batches an epoch visits and replay them through the training oracle.
Imported by tests/ only.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def round_keys(seed):
    """splitmix64 stream -> 8 round keys (bits 16..47 of each output)"""
    keys = []
    s = int(seed) & 0xFFFFFFFFFFFFFFFF
    for _ in range(8):
        s = (s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
        keys.append((z >> 16) & 0xFFFFFFFF)
    return keys


def _f(v, key):
    h = (v * np.uint64(0x9E3779B1) + np.uint64(key)) & M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE3D)) & M32
    h ^= h >> np.uint64(16)
    return h


def _once(x, keys, bits_lo, bits_hi):
    mlo = np.uint64((1 << bits_lo) - 1)
    mhi = np.uint64((1 << bits_hi) - 1)
    lo = x & mlo
    hi = x >> np.uint64(bits_lo)
    for r in range(0, 8, 2):
        hi = hi ^ (_f(lo, keys[r]) & mhi)
        lo = lo ^ (_f(hi, keys[r + 1]) & mlo)
    return (hi << np.uint64(bits_lo)) | lo


def epoch_positions(N, seed):
    """pos[r] for r in [0, N): a permutation of [0, N)."""
    N = int(N)
    bits = 2
    while (1 << bits) < N:
        bits += 1
    bits_lo, bits_hi = bits // 2, bits - bits // 2
    keys = round_keys(seed)
    x = _once(np.arange(N, dtype=np.uint64), keys, bits_lo, bits_hi)
    todo = np.nonzero(x >= N)[0]
    while todo.size:
        x[todo] = _once(x[todo], keys, bits_lo, bits_hi)
        todo = todo[x[todo] >= N]
    return x.astype(np.int64)


def epoch_batches(rec, pos, B):
    """stable multisplit: batch b = records with pos // B == b, in store order; batches concatenated"""
    batch = np.asarray(pos, np.int64) // int(B)
    order = np.argsort(batch, kind="stable")
    return rec[order], batch[order]
