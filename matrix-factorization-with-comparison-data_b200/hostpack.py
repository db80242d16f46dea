"""Host-side (numpy) packers for the staging formats of include/mfcd_b200.h.

A host that owns raw ``(u, i, j, z)`` records -- what the reference's DataLoader hands over, structure.py:846 --
can ship them as they are (16-byte records) or pack them first:

  pack8        8-byte hard-label records            (mfcd_unpack_triplets8 on the device)
  pack_wire    run-length format of ONE user-grouped batch, 4.375 B/triplet + 4 B/run
               (decoded by K1 itself, MFCD_FLAG_WIRE_RLE; mfcd_unpack_wire for other consumers)

These are the CPU counterparts of mfcd_pack_triplets8 / mfcd_pack_wire (bit-identical output, checked by the GPU
tests).  They run at numpy speed (~1e8 triplets/s per core) and are the readable statement of the formats; the
loaders themselves pack the 8-byte format with the library's C packer (mfcd_host_pack_triplets8, csrc/host_pack.cpp:
thread pool + AVX-512, ~3e9 triplets/s on 16 threads -- fast enough to pack every batch inside the step loop,
HostTripletLoader fmt="wire8_live").  pack_wire has no C twin: a host packs that format once and streams it every
epoch; bench.py reports the packing time next to the end-to-end numbers of the packed formats.
"""
from __future__ import annotations

import numpy as np


def as_records(u, i, j, z):
    """columns -> (N, 4) int32 records {u, i, j, float32 bits of z}"""
    rec = np.empty((len(u), 4), np.int32)
    rec[:, 0], rec[:, 1], rec[:, 2] = u, i, j
    rec[:, 3] = np.asarray(z, np.float32).view(np.int32)
    return rec


def pack8(rec):
    """(N, 4) int32 records -> uint64[N]: bit 0 label, [1,21) j, [21,41) i, [41,64) u."""
    rec = np.asarray(rec)
    z = rec[:, 3].view(np.float32)
    if not (((z == 0) | (z == 1)).all() and (rec[:, 0] >= 0).all() and (rec[:, 0] < (1 << 23)).all()
            and (rec[:, 1:3] >= 0).all() and (rec[:, 1:3] < (1 << 20)).all()):
        raise ValueError("pack8: soft labels or indices beyond 2^23 users / 2^20 items do not fit the 8-byte format")
    out = rec[:, 0].astype(np.uint64) << np.uint64(41)
    out |= rec[:, 1].astype(np.uint64) << np.uint64(21)
    out |= rec[:, 2].astype(np.uint64) << np.uint64(1)
    out |= (z != 0).astype(np.uint64)
    return out


def group_by_user(rec):
    """one batch with each user's records adjacent (stable), what the wire format and K1's user runs want"""
    rec = np.asarray(rec)
    return rec[np.argsort(rec[:, 0], kind="stable")]


def _bit_words(bits, nw):
    padded = np.zeros(nw * 32, np.uint8)
    padded[: len(bits)] = bits
    return np.packbits(padded.reshape(nw, 32), axis=1, bitorder="little").view(np.uint32).reshape(nw)


def pack_wire(rec):
    """ONE user-grouped batch of hard-labelled (B, 4) int32 records, items < 65536 -> uint32 words
    [n_runs, B, 0, 0 | word_run0[nw] | zbits[nw] | nbits[nw] | ij[B] | users[n_runs]], nw = ceil(B / 32)."""
    rec = np.asarray(rec)
    B = rec.shape[0]
    if B < 1:
        raise ValueError("pack_wire: empty batch")
    u, i, j = rec[:, 0], rec[:, 1], rec[:, 2]
    z = rec[:, 3].view(np.float32)
    if not (((z == 0) | (z == 1)).all() and (u >= 0).all() and (i >= 0).all() and (j >= 0).all()
            and (i < 65536).all() and (j < 65536).all()):
        raise ValueError("pack_wire: soft labels or item ids >= 65536 do not fit the run-length wire format")
    nw = (B + 31) // 32
    start = np.ones(B, np.uint8)
    start[1:] = u[1:] != u[:-1]
    before = np.cumsum(start, dtype=np.int64) - start          # run starts strictly before each triplet
    users = u[start.astype(bool)].astype(np.uint32)
    ij = i.astype(np.uint32) | (j.astype(np.uint32) << np.uint32(16))
    head = np.array([len(users), B, 0, 0], np.uint32)
    return np.concatenate([head, before[::32].astype(np.uint32), _bit_words((z != 0).astype(np.uint8), nw),
                           _bit_words(start, nw), ij, users])
