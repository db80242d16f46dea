"""The numpy restatement of the epoch batching (oracle/epoch_oracle.py): the keyed bijection really is a
permutation for every N, looks like a shuffle, and the stable multisplit forms exact batches.  (CPU only; the GPU
tests compare the kernels of csrc/epoch_batches.cu with this oracle bit for bit.)"""
import numpy as np
import pytest

from oracle import epoch_oracle as E


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 17, 255, 256, 257, 1000, 4097, 100_003])
def test_positions_are_a_permutation(N):
    p = E.epoch_positions(N, 12345 + N)
    assert p.dtype == np.int64 and np.array_equal(np.sort(p), np.arange(N))


def test_positions_behave_like_a_shuffle():
    N = 1 << 18
    r = np.arange(N)
    p = E.epoch_positions(N, 7)
    assert abs(np.corrcoef(r, p)[0, 1]) < 0.01
    assert abs(np.abs(p - r).mean() / N - 1 / 3) < 0.01                  # mean displacement of a uniform shuffle
    b = p // 4096
    assert abs((b[1:] == b[:-1]).mean() - 4096 / N) < 0.005              # neighbours land in the same batch by chance only
    q = E.epoch_positions(N, 8)
    assert abs(np.corrcoef(p, q)[0, 1]) < 0.01                           # epochs are independent
    assert np.array_equal(p, E.epoch_positions(N, 7))                    # and reproducible


def test_round_keys_known_answer():
    # splitmix64(seed = 0) first output is the published 0xE220A8397B1DCDAF; key = bits 16..47
    assert E.round_keys(0)[0] == (0xE220A8397B1DCDAF >> 16) & 0xFFFFFFFF


def test_batches_partition_the_store_in_store_order():
    rng = np.random.default_rng(0)
    N, B = 10_007, 1024
    rec = np.stack([np.sort(rng.integers(0, 50, N)), rng.integers(0, 9, N)], 1)
    pos = E.epoch_positions(N, 3)
    out, batch = E.epoch_batches(rec, pos, B)
    sizes = np.bincount(batch)
    assert sizes.tolist() == [B] * (N // B) + [N % B]
    for b in range(len(sizes)):
        rows = out[batch == b]
        assert (np.diff(rows[:, 0]) >= 0).all()                           # user-sorted store -> user-grouped batches
        members = np.nonzero(pos // B == b)[0]
        assert np.array_equal(rows, rec[members])                         # exactly the members, in store order
