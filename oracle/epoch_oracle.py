"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the epoch batching of csrc/epoch_batches.cu.

The reference reshuffles with torch's RandomSampler (structure.py:738); in device RNG mode the product path
replaces the materialised permutation by a keyed bijection of [0, N) (4 rounds of odd multiply / xor-shift / add
over ceil(log2 N) bits, with cycle walking) and forms batch b = {r : pos(r) // B == b}, each batch in store order.
This file restates both so that the tests can check the kernels bit for bit (index work), and so that a test can
rebuild the exact batches an epoch visits and replay them through the training oracle.
Imported by tests/ only.
"""
import numpy as np

MULS = (0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F)


def round_keys(seed, count=9):
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


def _once(x, keys, bits):
    """one pass of the bijection of the bits-wide integers: xor key, then 4 x (odd multiply, xor-shift, add key)"""
    mask = np.uint64((1 << bits) - 1)
    sh = np.uint64(bits >> 1)
    x = x ^ (np.uint64(keys[0]) & mask)
    for r in range(4):
        mul = np.uint64(((MULS[r] ^ ((keys[1 + r] << 1) & 0xFFFFFFFF)) | 1))
        x = (x * mul) & mask
        x = x ^ (x >> sh)
        x = (x + np.uint64(keys[5 + r])) & mask
    return x


def epoch_positions(N, seed):
    """pos[r] for r in [0, N): a permutation of [0, N)."""
    N = int(N)
    bits = 2
    while (1 << bits) < N:
        bits += 1
    keys = round_keys(seed)
    x = _once(np.arange(N, dtype=np.uint64), keys, bits)
    todo = np.nonzero(x >= N)[0]
    while todo.size:
        x[todo] = _once(x[todo], keys, bits)
        todo = todo[x[todo] >= N]
    return x.astype(np.int64)


def epoch_batches(rec, pos, B):
    """stable multisplit: batch b = records with pos // B == b, in store order; batches concatenated"""
    batch = np.asarray(pos, np.int64) // int(B)
    order = np.argsort(batch, kind="stable")
    return rec[order], batch[order]
