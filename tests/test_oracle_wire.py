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
