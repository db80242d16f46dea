"""The numpy restatement of the run-length staging format (oracle/wire_oracle.py): a hand-computed known-answer
vector for the layout in include/mfcd_b200.h, and round trips around word edges.  (CPU only; the GPU tests compare
the device packer / unpacker / K1 decoder against this oracle bit for bit.)"""
import numpy as np
import pytest

from oracle import wire_oracle as W


def test_known_answer_layout():
    # three triplets, users 7,7,9 -> runs start at 0 and 2; labels 1,0,1
    w = W.pack_wire([7, 7, 9], [1, 2, 3], [4, 5, 6], [1.0, 0.0, 1.0])
    expect = [2, 3, 0, 0,            # n_runs, B, 0, 0
              0,                     # word_run0[0]
              0b101,                 # zbits
              0b101,                 # nbits
              1 | 4 << 16, 2 | 5 << 16, 3 | 6 << 16,
              7, 9]                  # users
    assert w.dtype == np.uint32 and w.tolist() == expect


def test_second_word_starts_mid_run():
    # 40 triplets: users 0 x 35 then 5 x 5 -> word 1 begins inside the first run: word_run0 = [0, 1]
    u = np.array([0] * 35 + [5] * 5)
    i = np.arange(40); j = np.arange(40) + 100; z = (np.arange(40) % 3 == 0).astype(float)
    w = W.pack_wire(u, i, j, z)
    assert w[:4].tolist() == [2, 40, 0, 0] and w[4:6].tolist() == [0, 1]
    assert int(w[8]) == 1 and int(w[9]) == 1 << 3          # nbits: triplet 0, and triplet 35 = bit 3 of word 1
    assert w[-2:].tolist() == [0, 5]


@pytest.mark.parametrize("N,n", [(1, 5), (31, 3), (32, 40), (33, 2), (8191, 70), (8192, 1), (8193, 5000), (50_001, 999)])
def test_round_trip(N, n):
    rng = np.random.default_rng(N)
    u = np.sort(rng.integers(0, n, N))
    i, j = rng.integers(0, 65536, N), rng.integers(0, 65536, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    w = W.pack_wire(u, i, j, z)
    assert len(w) == 4 + 3 * ((N + 31) // 32) + N + len(np.unique(u))
    uu, ii, jj, zz = W.unpack_wire(w)
    assert (uu == u).all() and (ii == i).all() and (jj == j).all() and (zz == z).all()


@pytest.mark.parametrize("N,n", [(1, 5), (33, 2), (8193, 5000)])
def test_product_host_packers_match_the_oracle(N, n):
    """mfcd_b200.hostpack (the numpy packers a host-side loader uses) == the independent restatement above."""
    from mfcd_b200 import hostpack
    rng = np.random.default_rng(N + 1)
    u = np.sort(rng.integers(0, n, N))
    i, j = rng.integers(0, 65536, N), rng.integers(0, 65536, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    rec = hostpack.as_records(u, i, j, z)
    assert np.array_equal(hostpack.pack_wire(rec), W.pack_wire(u, i, j, z))
    p8 = hostpack.pack8(rec)
    assert np.array_equal(p8 >> np.uint64(41), u.astype(np.uint64))
    assert np.array_equal((p8 >> np.uint64(21)) & np.uint64(0xFFFFF), i.astype(np.uint64))
    assert np.array_equal((p8 >> np.uint64(1)) & np.uint64(0xFFFFF), j.astype(np.uint64))
    assert np.array_equal(p8 & np.uint64(1), z.astype(np.uint64))
    shuffled = rec[rng.permutation(N)]
    g = hostpack.group_by_user(shuffled)
    assert (np.diff(g[:, 0]) >= 0).all() and sorted(map(tuple, g)) == sorted(map(tuple, shuffled))
    soft = rec.copy(); soft[0, 3] = np.float32(0.5).view(np.int32)
    with pytest.raises(ValueError):
        hostpack.pack_wire(soft)
    with pytest.raises(ValueError):
        hostpack.pack8(soft)


def _pack8_by_the_book(u, i, j, z):
    """include/mfcd_b200.h: bit 0 = label, [1,21) = j, [21,41) = i, [41,64) = u -- restated with python ints"""
    return np.array([(int(a) << 41) | (int(b) << 21) | (int(c) << 1) | int(l != 0) for a, b, c, l in zip(u, i, j, z)],
                    dtype=np.uint64)


@pytest.mark.parametrize("N", [0, 1, 7, 8, 9, 63, 4097, (1 << 17) + 3])
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_c_host_packer_matches_the_layout(N, threads):
    """mfcd_host_pack_triplets8 (csrc/host_pack.cpp: thread pool + AVX-512 / scalar kernels; needs no GPU) against
    the header's bit layout: ragged sizes around the 8-record vector width, unaligned outputs, extreme indices,
    -0.0 labels, and the `bad` flag for soft labels / indices beyond the format."""
    import ctypes as C
    from mfcd_b200 import _lib, hostpack
    rng = np.random.default_rng(N + 17)
    u, i, j = rng.integers(0, 1 << 23, N), rng.integers(0, 1 << 20, N), rng.integers(0, 1 << 20, N)
    z = rng.integers(0, 2, N).astype(np.float64)
    if N > 2:
        u[0], i[0], j[0], z[0] = (1 << 23) - 1, (1 << 20) - 1, (1 << 20) - 1, 1.0
        u[1], i[1], j[1] = 0, 0, 0
    rec = hostpack.as_records(u, i, j, z)
    if N > 2:
        rec[1, 3] = np.float32(-0.0).view(np.int32)                         # -0.0 is label 0
        z[1] = 0.0
    want = _pack8_by_the_book(u, i, j, z)
    for shift in (0, 1):                                                     # 64-byte aligned or not
        buf = np.zeros(N + 9, np.uint64)
        off = (-(buf.ctypes.data // 8)) % 8 + shift
        out = buf[off:off + N]
        bad = C.c_int32(0)
        rc = _lib.lib.mfcd_host_pack_triplets8(rec.ctypes.data, N, out.ctypes.data, threads, C.byref(bad))
        assert rc == 0 and bad.value == 0
        assert np.array_equal(out, want)
        assert buf[:off].sum() == 0 and buf[off + N:].sum() == 0            # nothing written outside
    if N:
        assert np.array_equal(want, hostpack.pack8(rec))
    for col, val in ((0, 1 << 23), (0, -1), (1, 1 << 20), (2, 1 << 20), (3, int(np.float32(0.5).view(np.int32))),
                     (3, int(np.float32(2.0).view(np.int32)))):
        for pos in {0, N // 2, N - 1} if N else ():
            r = rec.copy(); r[pos, col] = val
            bad = C.c_int32(0)
            out = np.zeros(N, np.uint64)
            assert _lib.lib.mfcd_host_pack_triplets8(r.ctypes.data, N, out.ctypes.data, threads, C.byref(bad)) == 0
            assert bad.value == 1, (col, val, pos)
    assert _lib.lib.mfcd_host_pack_triplets8(None, 5, None, 1, None) == -1                 # MFCD_ERR_ARG


def test_c_host_packer_pool_under_concurrent_callers():
    """The packer's thread pool is shared by every caller in the process (the streaming epoch calls it from a worker
    thread while benchmarks / other loaders may call it too): concurrent calls with different thread counts must
    neither deadlock nor mix their chunks."""
    import ctypes as C
    import threading
    from mfcd_b200 import _lib, hostpack
    N = 150_001                                                   # > 2 chunks of 32768: the pool path
    recs, refs = [], []
    for seed in range(4):
        r = np.random.default_rng(seed)
        rec = hostpack.as_records(r.integers(0, 1 << 23, N), r.integers(0, 1 << 20, N), r.integers(0, 1 << 20, N),
                                  r.integers(0, 2, N).astype(np.float64))
        recs.append(rec)
        refs.append(hostpack.pack8(rec))
    errors = []

    def caller(w):
        out = np.zeros(N, np.uint64)
        bad = C.c_int32(0)
        for it in range(60):
            k, threads = (w + it) % 4, (1, 2, 5, 16, 0)[(it + w) % 5]
            rc = _lib.lib.mfcd_host_pack_triplets8(recs[k].ctypes.data, N, out.ctypes.data, threads, C.byref(bad))
            if rc or bad.value or not np.array_equal(out, refs[k]):
                errors.append((w, it, rc, bad.value))
                return

    ts = [threading.Thread(target=caller, args=(w,), daemon=True) for w in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "packer pool deadlocked"
    assert not errors, errors[:3]


def test_c_host_packer_scalar_kernel_in_a_fresh_process():
    """Hosts without AVX-512 take the portable loop; force it (MFCD_HOST_PACK_ISA=scalar is read once per process)
    and compare whole batches, not just the vector kernel's tail, with the numpy packer."""
    import os
    import subprocess
    import sys
    code = r"""
import ctypes as C, numpy as np, sys
sys.path.insert(0, %r)
import mfcd_b200
from mfcd_b200 import _lib, hostpack
assert _lib.lib.mfcd_host_pack_isa() == 0, "scalar kernel was not selected"
rng = np.random.default_rng(5)
N = 100_003
rec = hostpack.as_records(rng.integers(0, 1 << 23, N), rng.integers(0, 1 << 20, N), rng.integers(0, 1 << 20, N),
                          rng.integers(0, 2, N).astype(np.float64))
for threads in (1, 4):
    out = np.zeros(N, np.uint64); bad = C.c_int32(0)
    assert _lib.lib.mfcd_host_pack_triplets8(rec.ctypes.data, N, out.ctypes.data, threads, C.byref(bad)) == 0
    assert bad.value == 0 and np.array_equal(out, hostpack.pack8(rec))
rec[77, 3] = np.float32(0.25).view(np.int32)
assert _lib.lib.mfcd_host_pack_triplets8(rec.ctypes.data, N, out.ctypes.data, 4, C.byref(bad)) == 0 and bad.value == 1
print("scalar ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MFCD_HOST_PACK_ISA="scalar")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "scalar ok" in r.stdout, r.stdout + r.stderr
