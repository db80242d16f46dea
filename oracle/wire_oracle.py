"""CPU restatement of the run-length staging format (TEST INFRASTRUCTURE ONLY -- imported by tests/, never by
the product path).  The format is this repo's own (the reference moves python tuples: structure.py:846 does
`x.to(device)` on int64/float64 batches), so there is no reference golden for it; what is pinned here is the
layout documented in include/mfcd_b200.h against an independent numpy implementation:

    u32 words  [n_runs, B, 0, 0 | word_run0[nw] | zbits[nw] | nbits[nw] | ij[B] | users[n_runs]],  nw = ceil(B/32)
    nbits: bit p of word w = 1 iff triplet 32w+p opens a run of equal users (triplet 0 always does)
    word_run0[w] = number of run starts before triplet 32w;  zbits = label bits;  ij = i | j << 16
"""
import numpy as np


def pack_wire(u, i, j, z):
    """(u, i, j, z) columns of ONE user-grouped batch (hard labels, items < 65536) -> uint32 words."""
    u = np.asarray(u, np.int64); i = np.asarray(i, np.int64); j = np.asarray(j, np.int64)
    z = np.asarray(z, np.float64)
    B = len(u)
    assert B >= 1 and ((z == 0) | (z == 1)).all() and i.max() < 65536 and j.max() < 65536 and i.min() >= 0 and j.min() >= 0
    nw = (B + 31) // 32
    start = np.ones(B, bool)
    start[1:] = u[1:] != u[:-1]
    pad = nw * 32 - B
    bit_weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))

    def words(bits):
        b = np.concatenate([bits, np.zeros(pad, bool)]).reshape(nw, 32).astype(np.uint64)
        return (b * bit_weights).sum(axis=1).astype(np.uint32)

    starts_before = np.concatenate([[0], np.cumsum(start)])[:-1]         # run starts strictly before triplet k
    word_run0 = starts_before[::32].astype(np.uint32)
    users = u[start].astype(np.uint32)
    ij = (i.astype(np.uint32) | (j.astype(np.uint32) << np.uint32(16)))
    head = np.array([len(users), B, 0, 0], np.uint32)
    return np.concatenate([head, word_run0, words(z != 0), words(start), ij, users])


def unpack_wire(wire):
    """uint32 words -> (u, i, j, z) int64/int64/int64/float64 columns."""
    wire = np.asarray(wire, np.uint32)
    n_runs, B = int(wire[0]), int(wire[1])
    nw = (B + 31) // 32
    o = 4
    word_run0 = wire[o:o + nw]; o += nw
    zbits = wire[o:o + nw]; o += nw
    nbits = wire[o:o + nw]; o += nw
    ij = wire[o:o + B]; o += B
    users = wire[o:o + n_runs]
    k = np.arange(B)
    w, p = k // 32, k % 32
    upto = (np.uint64(0xFFFFFFFF) >> (np.uint64(31) - p.astype(np.uint64))).astype(np.uint64)   # bits 0..p
    masked = nbits[w].astype(np.uint64) & upto
    pop = np.array([bin(int(x)).count("1") for x in masked], np.int64)
    run = word_run0[w].astype(np.int64) + pop - 1
    u = users[run].astype(np.int64)
    i = (ij & np.uint32(0xFFFF)).astype(np.int64)
    j = (ij >> np.uint32(16)).astype(np.int64)
    z = ((zbits[w] >> p.astype(np.uint32)) & np.uint32(1)).astype(np.float64)
    return u, i, j, z
