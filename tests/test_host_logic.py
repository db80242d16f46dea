"""Host-side logic that needs no GPU: sweep expansion, reference-RNG replay, plans."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import load_golden

import structure
import generation_data
from mfcd_b200 import config, sampling
from mfcd_b200 import dist as mdist
from mfcd_b200.trainer import OptimizerSpec, resolve_mode, MatrixFactorization


def test_normalise_grid_and_numpy_scalars():
    grid, lists, sync = structure._normalise_grid({"a": np.float64(0.5), "b": [np.int64(1), 2], "c": np.array([3, 4])})
    assert grid["a"] == [0.5] and type(grid["a"][0]) is float
    assert grid["b"] == [1, 2] and type(grid["b"][0]) is int
    assert len(lists) == 2 and sync
    _, _, sync = structure._normalise_grid({"a": [1, 2, 3], "b": [1, 2]})
    assert not sync


def _fake_run(calls):
    def run_experiment(**kw):
        calls.append(kw)
        return {"accuracy": [0.5]}
    return run_experiment


def test_parameter_scan_grid_linear_and_errors(monkeypatch, tmp_path, capsys):
    calls = []
    monkeypatch.setattr(structure, "run_experiment", _fake_run(calls))
    res = structure.parameter_scan(n=10, m=10, d=[2, 4], p=[0.1, 0.2, 0.3], num_epochs=1)
    assert len(res) == 6 and len(calls) == 6
    assert [r["params"]["d"] for r in res] == [2, 2, 2, 4, 4, 4]          # itertools.product order over the key order
    assert set(res[0]["params"]) == set(structure._SCAN_KEYS) and len(structure._SCAN_KEYS) == 16
    calls.clear()
    res = structure.parameter_scan(n=10, m=10, p=[0.1, 0.2], K=[1, 3], linear=True)
    assert [(c["p"], c["K"]) for c in calls] == [(0.1, 1), (0.2, 3)]
    with pytest.raises(ValueError, match="not synchronized"):
        structure.parameter_scan(n=10, m=10, p=[0.1, 0.2], K=[1, 2, 3], linear=True)


def test_parameter_scan_saving_semantics(monkeypatch, tmp_path):
    calls = []
    monkeypatch.setattr(structure, "run_experiment", _fake_run(calls))
    path = str(tmp_path / "out" / "scan.pkl")
    os.makedirs(os.path.dirname(path))
    with open(path, "wb") as f:
        pickle.dump(["stale"], f)
    res = structure.parameter_scan(n=10, m=10, p=[0.1, 0.2, 0.3], save_path=path, save_every=2)
    assert res == []                                   # saved chunks are dropped from the return value
    with open(path, "rb") as f:
        stored = pickle.load(f)
    assert len(stored) == 3 and "stale" not in stored  # an existing file is deleted first
    with pytest.raises(TypeError):
        structure.parameter_scan(n=10, filename="x")   # Runs.ipynb cell 13's bug stays a TypeError


def test_unknown_names_raise_like_the_reference():
    X = torch.zeros(5, 5)
    with pytest.raises(ValueError, match="Unknown triplet sampling strategy"):
        structure.get_triplets_from_X(X, 3, strategy="nope")
    with pytest.raises(ValueError, match="Unknown generation method"):
        structure.generate_X(5, 5, 2, "cpu", generation="nope")
    with pytest.raises(NotImplementedError):
        structure.generate_X(5, 5, 2, "cpu", generation="gmm")
    with pytest.raises(NotImplementedError):
        structure.get_triplets_from_X(X, 3, strategy="cluster")


def test_reference_rng_replay_random_sampler():
    """host_sample_random == the reference's choose_items_random under the same torch seed
    (golden 'random' was recorded from the reference with seed 3), order included."""
    g = load_golden("samplers.npz")
    torch.manual_seed(3)
    mine = sampling.host_sample_random(40, 30, 200, set())
    assert np.array_equal(np.array(mine, np.int64), g["random"])
    # generator left exactly where a one-at-a-time loop leaves it: replay twice -> same continuation
    torch.manual_seed(3)
    sampling.host_sample_random(40, 30, 200, set())
    a = torch.rand(4)
    torch.manual_seed(3)
    sampling.host_sample_random(40, 30, 200, set())
    assert torch.equal(a, torch.rand(4))


def test_reference_rng_replay_popularity(monkeypatch):
    g = load_golden("samplers.npz")
    monkeypatch.setattr(config, "RNG_MODE", "reference")
    X = torch.from_numpy(g["X"])
    torch.manual_seed(4); np.random.seed(4)
    mine = generation_data.choose_items_by_popularity(X, 150, set(), method="zipf", alpha=1.5)
    assert np.array_equal(np.array(mine, np.int64), g["popularity_zipf"])
    torch.manual_seed(4); np.random.seed(4)
    mine = generation_data.choose_items_by_popularity(X, 100, set(), method="exponential", alpha=0.2)
    assert np.array_equal(np.array(mine, np.int64), g["popularity_exp"])
    with pytest.raises(ValueError, match="Unknown popularity method"):
        generation_data.choose_items_by_popularity(X, 5, set(), method="nope")


def test_reference_generate_x_matches_golden(monkeypatch):
    g = load_golden("train_c1.npz")
    monkeypatch.setattr(config, "RNG_MODE", "reference")
    torch.manual_seed(0); np.random.seed(0)
    X = structure.generate_X(100, 100, 2, "cpu")
    assert np.array_equal(X.numpy(), g["X"])


def test_model_init_matches_reference_stream():
    g = load_golden("train_c1.npz")
    # golden: seed 0, X (numpy RNG only), split (torch RNG), then the model init
    U0 = g["U0"]
    torch.manual_seed(123)
    model = MatrixFactorization(7, 5, 4)
    torch.manual_seed(123)
    U = torch.randn(7, 4) / torch.sqrt(torch.tensor(4, dtype=torch.float32))
    V = torch.randn(5, 4) / torch.sqrt(torch.tensor(4, dtype=torch.float32))
    assert torch.equal(model.U.detach(), U) and torch.equal(model.V.detach(), V)
    assert U0.shape == (100, 2)


def test_optimizer_spec_and_modes():
    p = [torch.nn.Parameter(torch.zeros(3))]
    s = OptimizerSpec(torch.optim.Adam(p, lr=2e-3, weight_decay=1e-5))
    assert (s.kind, s.lr, s.weight_decay, s.beta1, s.beta2, s.eps) == (0, 2e-3, 1e-5, 0.9, 0.999, 1e-8)
    s = OptimizerSpec(torch.optim.SGD(p, lr=0.1, momentum=0.9))
    assert (s.kind, s.momentum) == (1, 0.9)
    with pytest.raises(NotImplementedError):
        OptimizerSpec(torch.optim.Adam(p, amsgrad=True))
    with pytest.raises(NotImplementedError):
        OptimizerSpec(torch.optim.RMSprop(p))
    assert resolve_mode("auto", 64) == 1 and resolve_mode("auto", 1 << 20) == 0
    assert resolve_mode("atomic", 64) == 0 and resolve_mode("deterministic", 1 << 20) == 1
    with pytest.raises(ValueError):
        resolve_mode("nope", 64)


def test_plans_cover_every_sample_once():
    N, B, W = 1003, 96, 4
    seen = np.zeros(N, int)
    for r in range(W):
        plan = mdist.ReplicatedPlan(N, B, r, W)
        for k in range(plan.n_steps()):
            s, bl, bg = plan.local_range(k)
            assert bg == min(B, N - k * B)
            seen[s:s + bl] += 1
    assert (seen == 1).all()
    sizes = [250, 251, 249, 253]
    plans = [mdist.PartitionedPlan(sizes, 64, r) for r in range(4)]
    assert all(p.n_steps() == 4 for p in plans)
    for k in range(4):
        tot = sum(p.local_range(k)[1] for p in plans)
        assert all(p.local_range(k)[2] == tot for p in plans)
    assert sum(p.local_range(k)[1] for p in plans for k in range(4)) == sum(sizes)
    assert mdist.bucket_bounds(10, 4) == [(0, 4), (4, 8), (8, 10)]


def test_popularity_cdf_and_keys():
    cdf = sampling.popularity_cdf(50, "zipf", 1.5)
    assert abs(cdf[-1] - 1) < 1e-12 and (np.diff(cdf) > 0).all()
    with pytest.raises(ValueError):
        sampling.popularity_cdf(5, "nope")
    assert structure._split_permutation(10, None).tolist() == \
        torch.randperm(10, generator=torch.Generator().manual_seed(42)).tolist()


@pytest.mark.skipif(not os.path.isfile("/root/reference/generation_data.py"), reason="no reference checkout here")
def test_out_of_scope_names_pass_through_to_an_importable_reference(monkeypatch):
    """cluster / user_similarity and the ten other generators run the reference's own host code when
    MFCD_REFERENCE_PATH names a checkout (Runs.ipynb cell 16 sweeps "cluster"); without it they raise."""
    import generation_data as gen
    monkeypatch.setenv("MFCD_REFERENCE_PATH", "/root/reference")
    monkeypatch.setattr(gen, "_REFERENCE_MODULE", None)
    torch.manual_seed(0); np.random.seed(0)
    X = torch.randn(12, 30)
    got = structure.get_triplets_from_X(X, 25, strategy="cluster", n_clusters=4)
    assert isinstance(got, set) and len(got) == 25
    assert all(0 <= u < 12 and 0 <= i < 30 and 0 <= j < 30 and i != j for u, i, j in got)
    Xg = structure.generate_X(9, 11, 3, "cpu", generation="gmm")
    assert tuple(Xg.shape) == (9, 11)
    Xl = structure.generate_X(9, 11, 3, "cpu", generation="low_rank")
    assert tuple(Xl.shape) == (9, 11)
    monkeypatch.setattr(gen, "_REFERENCE_MODULE", None)
    monkeypatch.delenv("MFCD_REFERENCE_PATH")
    with pytest.raises(NotImplementedError):
        structure.get_triplets_from_X(X, 3, strategy="cluster")
    monkeypatch.setattr(gen, "_REFERENCE_MODULE", None)
