"""CUDA matching kernels vs the oracle / the reference's golden outputs, through the C ABI."""
import numpy as np
import pytest
import torch

from lidarregistration_b200 import engine, synthetic
from lidarregistration_b200.algorithms import matching as GM
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu


def test_golden_find_nn_and_mutual(matching_golden):
    for name, c in matching_golden.items():
        f0, f1 = torch.from_numpy(c["f0"]), torch.from_numpy(c["f1"])
        i0, i1, i2 = GM.find_nn(f0, f1, return_2nd=True)
        assert i1.dtype == torch.int64 and not i1.is_cuda
        assert np.array_equal(i1.numpy(), c["idx1"]), name
        assert np.array_equal(i2.numpy(), c["idx2"]), name
        assert np.array_equal(i0.numpy(), np.arange(len(c["f0"])))
        _, j1, none = GM.find_nn(f0, f1)
        assert none is None and torch.equal(j1, i1)
        m0, m1, m2 = GM.nn_to_mutual(f0, f1, i0, i1, i2)
        assert np.array_equal(m0.numpy(), c["mut_i"]) and np.array_equal(m1.numpy(), c["mut_j"]), name
        assert np.array_equal(m2.numpy(), c["mut_2nd"]), name
        r = GM.calc_distance_ratio_in_feature_space(f0, f1, m0, m1, m2).numpy()
        ok = np.isfinite(c["ratio"])
        assert np.allclose(r[ok], c["ratio"][ok], rtol=3e-7, atol=0), name
        ro = O.ratio(c["f0"], c["f1"], c["mut_i"], c["mut_j"], c["mut_2nd"])
        assert np.array_equal(r[ok], ro[ok]), name  # bit-exact against the oracle


@pytest.mark.parametrize("N,M,seed", [(5000, 6000, 1), (4097, 3001, 2), (20000, 20000, 3)])
def test_find_nn_vs_oracle_synthetic(N, M, seed):
    p = synthetic.make_pair(N, M, seed=seed, overlap=0.6)
    i1, i2 = engine.match_nn(p["feat0"], p["feat1"], want_2nd=True)
    _, o1, o2 = O.find_nn(p["feat0"], p["feat1"], return_2nd=True)
    assert np.array_equal(i1.cpu().numpy(), o1)
    assert np.array_equal(i2.cpu().numpy(), o2)
    mi, mj = engine.match_mutual(p["feat0"], p["feat1"], i1)
    oi, oj = O.nn_to_mutual(p["feat0"], p["feat1"], o1)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mj.cpu().numpy(), oj)


def test_find_2nn_contract():
    p = synthetic.make_pair(3000, seed=8)
    i0, i1, i2, extra = GM.find_2nn(torch.from_numpy(p["feat0"]), torch.from_numpy(p["feat1"]))
    assert isinstance(extra, float) and len(i0) == len(i1) == len(i2) == 3000
    assert not torch.any(i1 == i2)


def test_full_size_properties():
    """cfg 2 of BASELINE.json (N = M = 50k x 32): size-independent properties instead of the oracle."""
    N = M = 50000
    g = torch.Generator(device="cuda").manual_seed(5)
    f0 = torch.nn.functional.normalize(torch.randn(N, 32, device="cuda", generator=g), dim=1)
    f1 = torch.nn.functional.normalize(torch.randn(M, 32, device="cuda", generator=g), dim=1)
    f1[:20000] = torch.nn.functional.normalize(f0[:20000] + 0.05 * torch.randn(20000, 32, device="cuda", generator=g), dim=1)
    i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
    assert int(i1.min()) >= 0 and int(i1.max()) < M and not torch.any(i1 == i2)
    # planted neighbours are found
    assert (i1[:20000] == torch.arange(20000, device="cuda")).float().mean() > 0.99
    # d(i, nn) <= d(i, 2nd) <= d(i, random j)
    d = lambda a, b: (a - b).norm(dim=1)
    rnd = torch.randint(0, M, (N,), device="cuda")
    assert torch.all(d(f0, f1[i1]) <= d(f0, f1[i2]) + 1e-6)
    assert torch.all(d(f0, f1[i2]) <= d(f0, f1[rnd]) + 1e-6 + (rnd == i1) * 10)
    # mutual is an involution: matching the other way gives the transposed pair set
    mi, mj = engine.match_mutual(f0, f1, i1)
    r1, _ = engine.match_nn(f1, f0)
    ni, nj = engine.match_mutual(f1, f0, r1)
    a = torch.stack([mi, mj], 1)
    b = torch.stack([nj, ni], 1)
    b = b[torch.argsort(b[:, 0])]
    assert torch.equal(a, b) and torch.all(mi[1:] > mi[:-1])
    # EVERY row against the oracle arithmetic (cfg 2 at full size: the oracle's 50k x 50k sweep takes seconds on the
    # host cores), both neighbours, and the mutual set
    _, o1, o2 = O.find_nn(f0.cpu().numpy(), f1.cpu().numpy(), return_2nd=True)
    assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2)
    q0, q1 = O.nn_to_mutual(f0.cpu().numpy(), f1.cpu().numpy(), o1)
    assert np.array_equal(mi.cpu().numpy(), q0) and np.array_equal(mj.cpu().numpy(), q1)


def test_gather_and_errors():
    xyz = torch.arange(30, dtype=torch.float32).reshape(10, 3)
    idx = torch.tensor([9, 0, 3, 3])
    out = engine.gather_xyz(xyz, idx).cpu()
    assert torch.equal(out, xyz[idx])
    with pytest.raises(RuntimeError):
        engine.match_nn(torch.zeros(4, 20), torch.zeros(4, 20))  # unsupported D


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_both_sweep_implementations_agree_with_oracle(mode):
    """mode 0 = default, 1 = exact CUDA-core sweep, 2 / 3 = tcgen05 sweep with fp32 / fp16 accumulators
    + exact re-rank: identical indices."""
    engine.match_set_mode(mode)
    try:
        for (N, M, seed) in [(1000, 777, 1), (300, 5000, 2), (2500, 260, 3)]:
            rng = np.random.default_rng(seed)
            f0 = rng.standard_normal((N, 32)).astype(np.float32) * rng.uniform(0.5, 2.0, (N, 1)).astype(np.float32)
            f1 = rng.standard_normal((M, 32)).astype(np.float32) * rng.uniform(0.5, 2.0, (M, 1)).astype(np.float32)
            k = min(N, M) // 2
            f1[:k] = f0[:k] + 0.05 * rng.standard_normal((k, 32)).astype(np.float32)
            i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
            j1, _ = engine.match_nn(f0, f1, want_2nd=False)
            _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
            assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2), (mode, N, M)
            assert torch.equal(i1, j1)
            mi, mj = engine.match_mutual(f0, f1, i1)
            oi, oj = O.nn_to_mutual(f0, f1, o1)
            assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mj.cpu().numpy(), oj), (mode, N, M)
    finally:
        engine.match_set_mode(0)


@pytest.mark.parametrize("mode", [2, 3])
def test_candidate_overflow_falls_back_to_exact_scan(mode):
    """More near-ties than candidate slots: the tensor-core path must hand the row to the exact scan."""
    engine.match_set_mode(mode)
    try:
        _overflow_case()
    finally:
        engine.match_set_mode(0)


def _overflow_case():
    rng = np.random.default_rng(9)
    f1 = rng.standard_normal((3000, 32)).astype(np.float32)
    f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
    f1[500:700] = f1[500] + 1e-4 * rng.standard_normal((200, 32)).astype(np.float32)  # 200 near-duplicates
    f1[1000:1100] = f1[1000]                                                            # 100 exact duplicates
    f0 = np.concatenate([f1[500:520], f1[1000:1010], f1[:300]]).copy()
    i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
    _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
    assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2)
    assert np.all(o1[20:30] == 1000) and np.all(o2[20:30] == 1001)


@pytest.mark.parametrize("mode", [2, 3])
def test_event_overflow_falls_back_to_exact_scan(mode):
    """Targets that get steadily closer to the query along the column order make every 32-column chunk a new
    running-maximum record: more events than slots, so the tensor-core path must hand the rows to the exact scan."""
    rng = np.random.default_rng(11)
    M = 40000
    base = rng.standard_normal((8, 32)).astype(np.float32)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    f1 = rng.standard_normal((M, 32)).astype(np.float32)
    f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
    w = np.linspace(0.0, 1.0, M, dtype=np.float32)[:, None]
    f1 = (1.0 - w) * f1 + w * base[0]          # approaches base[0] monotonically
    f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
    f0 = np.concatenate([base, f1[::997][:40]]).copy()
    engine.match_set_mode(mode)
    try:
        i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
        j1, _ = engine.match_nn(f0, f1, want_2nd=False)
    finally:
        engine.match_set_mode(0)
    _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
    assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2)
    assert np.array_equal(j1.cpu().numpy(), o1)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_tiny_shapes(mode):
    engine.match_set_mode(mode)
    try:
        rng = np.random.default_rng(1)
        for (N, M) in [(1, 1), (1, 5), (7, 1), (2, 2), (128, 256), (129, 257)]:
            f0 = rng.standard_normal((N, 32)).astype(np.float32)
            f1 = rng.standard_normal((M, 32)).astype(np.float32)
            i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
            _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
            assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2), (N, M)
            mi, mj = engine.match_mutual(f0, f1, i1)
            oi, oj = O.nn_to_mutual(f0, f1, o1)
            assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mj.cpu().numpy(), oj), (N, M)
    finally:
        engine.match_set_mode(0)


@pytest.mark.parametrize("mode", [2, 3])
def test_uniform_norm_path_with_negative_scores_and_padding(mode):
    """Unit-norm features take the K = 32 path (no norm-extension MMA), where padding rows would score v = 0.
    Queries anti-correlated with every target have only negative scores: the padded columns of the last tile
    must not win or raise thresholds.  M is chosen so that the last tile is mostly padding."""
    rng = np.random.default_rng(21)
    for M in (257, 300, 1000, 5121):
        f1 = np.abs(rng.standard_normal((M, 32))).astype(np.float32) + 0.1   # all-positive targets
        f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
        f0 = -np.abs(rng.standard_normal((700, 32))).astype(np.float32) - 0.1  # all-negative queries
        f0 /= np.linalg.norm(f0, axis=1, keepdims=True)
        engine.match_set_mode(mode)
        try:
            i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
            r1, _ = engine.match_nn(f1, f0)
        finally:
            engine.match_set_mode(0)
        _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
        assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2), M
        _, p1, _ = O.find_nn(f1, f0)
        assert np.array_equal(r1.cpu().numpy(), p1), M


@pytest.mark.parametrize("mode", [2, 3])
def test_dense_near_ties_within_the_band(mode):
    """Clusters of targets a few 1e-4 apart (inside the fp16 band): many events and flagged sub-groups per row,
    first and second neighbour often in the same 32-column chunk; indices must still be the oracle's."""
    rng = np.random.default_rng(22)
    centres = rng.standard_normal((400, 32)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    f1 = np.repeat(centres, 12, axis=0) + 2e-4 * rng.standard_normal((4800, 32)).astype(np.float32)
    f1 = f1[rng.permutation(len(f1))]
    f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
    f0 = centres + 1e-4 * rng.standard_normal(centres.shape).astype(np.float32)
    f0 /= np.linalg.norm(f0, axis=1, keepdims=True)
    engine.match_set_mode(mode)
    try:
        i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
    finally:
        engine.match_set_mode(0)
    _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
    assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2)
