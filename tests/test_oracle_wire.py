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
