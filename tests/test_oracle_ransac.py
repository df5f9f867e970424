"""Known-answer tests pinning the oracle's RANSAC half (the reference ships no fixture for it:
SURVEY.md 8(c) "parity unpinned"), plus the reference's own Kabsch witness
(Experiments/models/common.py:7-45) through tests/golden/kabsch_ref.npz."""
import numpy as np
import pytest

from lidarregistration_b200 import synthetic
from oracle import lr_oracle as O


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def cost(T, P, Q):
    return float(np.sum((P @ T[:3, :3].T + T[:3, 3] - Q) ** 2))


def test_kabsch_against_reference_witness(kabsch_golden):
    g = kabsch_golden
    for P, Q, T, k in zip(g["P"], g["Q"], g["T"], g["k"]):
        P, Q = P[:k].astype(np.float64), Q[:k].astype(np.float64)
        mine = O.kabsch(P, Q)
        assert abs(np.linalg.det(mine[:3, :3]) - 1) < 1e-12
        assert np.allclose(mine[:3, :3] @ mine[:3, :3].T, np.eye(3), atol=1e-12)
        # least squares: never worse than the reference's fp32 solution
        assert cost(mine, P, Q) <= cost(T.astype(np.float64), P, Q) * (1 + 1e-4) + 1e-6
        if cost(mine, P, Q) < 0.1 * k:  # related clouds: the optimum is well conditioned
            assert np.allclose(mine, T, atol=2e-3)


def test_kabsch_closed_form_cases():
    rng = np.random.default_rng(0)
    P = rng.uniform(-50, 50, (3, 3))
    assert np.allclose(O.kabsch(P, P), np.eye(4), atol=1e-12)                      # identity
    t = np.array([3.0, -2.0, 0.5])
    T = O.kabsch(P, P + t)                                                           # pure translation
    assert np.allclose(T[:3, :3], np.eye(3), atol=1e-12) and np.allclose(T[:3, 3], t, atol=1e-10)
    R = rot([0, 0, 1], np.pi)                                                        # 180 deg yaw
    T = O.kabsch(P, P @ R.T)
    assert np.allclose(T[:3, :3], R, atol=1e-10)
    # reflection: the best proper rotation is returned, never det = -1
    Q = P * np.array([1.0, 1.0, -1.0])
    T = O.kabsch(np.vstack([P, rng.uniform(-50, 50, (5, 3))]), np.vstack([Q, rng.uniform(-50, 50, (5, 3))]))
    assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-12
    # collinear sample (rank 1) and coincident points (rank 0) stay finite and orthonormal
    L = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0]], float)
    for Pc, Qc in ((L, L @ rot([0, 1, 0], 0.3).T + 1.0), (np.zeros((3, 3)), np.ones((3, 3)))):
        T = O.kabsch(Pc, Qc)
        assert np.all(np.isfinite(T)) and np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)
    assert np.allclose(O.kabsch(np.zeros((0, 3)), np.zeros((0, 3))), np.eye(4))     # empty -> identity


def test_elc_predicate():
    # preemption_edge_length.h:116-123: fail iff ds < 0.9 dt or dt < 0.9 ds for any pair
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    assert O.elc(P, P * 1.0)
    assert O.elc(P, P * 1.1)            # 1/1.1 = 0.909 > 0.9
    assert not O.elc(P, P * 1.12)       # 1/1.12 = 0.893 < 0.9
    assert not O.elc(P * 1.12, P)
    Q = P.copy(); Q[2] = [0, 5, 0]
    assert not O.elc(P, Q)              # one bad edge is enough
    D = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0]], float)
    assert O.elc(D, D)                  # duplicate index: 0 < 0 is false -> passes (App. B)
    assert O.elc(np.vstack([P, [[1, 1, 1]]]), np.vstack([P, [[1, 1, 1]]]))  # m = 4, 6 edges


def test_sampler_properties():
    n = 37
    seen = set()
    for hid in range(2000):
        s = O.sample(51, hid, 0, 3, n)
        assert len(set(s.tolist())) == 3 and s.min() >= 0 and s.max() < n
        seen.update(s.tolist())
        s4 = O.sample(51, hid, 0, 4, n)
        assert len(set(s4.tolist())) == 4
        r = O.sample(51, hid, 2, 4, n)
        assert r.min() >= 0 and r.max() < n
    assert seen == set(range(n))
    assert np.array_equal(O.sample(7, 123, 0, 3, 1000), O.sample(7, 123, 0, 3, 1000))
    assert not np.array_equal(O.sample(7, 123, 0, 3, 1000), O.sample(8, 123, 0, 3, 1000))
    # n == m: a permutation
    assert sorted(O.sample(1, 5, 0, 3, 3).tolist()) == [0, 1, 2]


def test_prosac_sampler_properties():
    # draw k uses the n_k best correspondences, the newest one is forced, uniform after T_N
    n, m = 5000, 3
    g = O.prosac_growth(n, m)
    assert np.all(g[:m] == 1) and np.all(np.diff(g[m - 1:].astype(np.int64)) >= 1)
    prev_top = 0
    for hid in list(range(0, 300)) + [5000, 50000, 99999]:
        s = O.sample(9, hid, O.PROSAC, m, n, g)
        assert len(set(s.tolist())) == m and s.max() < n
        k = hid + 1
        nk = m + int(np.searchsorted(g[m - 1:], k, side="left"))  # smallest n >= m with g[n-1] >= k
        assert s[-1] == nk - 1 and s[:-1].max() < nk - 1
        assert nk >= prev_top
        prev_top = nk
    assert O.sample(9, 0, O.PROSAC, m, n, g).tolist()[-1] == m - 1      # first draw: the m best
    late = np.array([O.sample(9, h, O.PROSAC, m, n, g) for h in range(100000, 100400)])
    assert np.array_equal(late, np.array([O.sample(9, h, O.UNIFORM, m, n) for h in range(100000, 100400)]))
    # sorted-by-quality inputs: PROSAC finds the model in far fewer draws than uniform sampling
    d = synthetic.make_correspondences(4000, inlier_ratio=0.1, seed=3)
    order = np.argsort(~d["is_inlier"], kind="stable")                   # inliers first = perfect quality ranking
    src, tgt = d["src"][order], d["tgt"][order]
    a = O.ransac(src, tgt, sampler=O.PROSAC, conf=1.0, max_iters=64, round_size=64, seed=1)
    b = O.ransac(src, tgt, sampler=O.UNIFORM, conf=1.0, max_iters=64, round_size=64, seed=1)
    assert a["best_count"] > 300 > b["best_count"]


def test_planted_inliers_zero_noise():
    d = synthetic.make_correspondences(2000, inlier_ratio=0.3, seed=77, noise=0.0)
    T = d["T_gt"]
    cnt, mask = O.count_inliers(d["src"], d["tgt"], T, 0.6, return_mask=True)
    assert cnt >= d["is_inlier"].sum() and np.all(mask[d["is_inlier"]])
    res = O.ransac(d["src"], d["tgt"], m=3, sampler=0, use_elc=True, thr=0.6, conf=1.0, max_iters=3000, seed=51)
    assert res["best_count"] >= d["is_inlier"].sum()
    assert np.allclose(res["T_refit"], T, atol=1e-4)


def test_fed_samples_and_selection():
    d = synthetic.make_correspondences(1500, inlier_ratio=0.4, seed=5)
    rng = np.random.default_rng(1)
    samples = rng.integers(0, 1500, (400, 3)).astype(np.int32)
    counts, best, models = O.score_samples(d["src"], d["tgt"], samples, 0.6, True, 0.9, return_models=True)
    assert best == int(np.argmax(counts))  # first maximum
    for h in (0, 17, best):
        P, Q = d["src"][samples[h]].astype(float), d["tgt"][samples[h]].astype(float)
        if counts[h] < 0:
            assert not O.elc(P, Q)
        else:
            T = np.eye(4); T[:3] = models[h].reshape(3, 4)
            assert np.allclose(T, O.kabsch(P, Q), atol=0) and counts[h] == O.count_inliers(d["src"], d["tgt"], T, 0.6)
    none, nb = O.score_samples(d["src"], d["tgt"][::-1].copy(), samples[:5] * 0, 0.6, False)
    assert nb == 0  # ELC off: degenerate (repeated index) samples still get a model


def test_confidence_rule_and_round_granularity():
    assert O.conf_iters(9000, 30000, 3, 0.9995, 10**6) == 278
    assert O.conf_iters(0, 30000, 3, 0.9995, 10**6) == 10**6
    assert O.conf_iters(30000, 30000, 3, 0.9995, 10**6) == 1
    assert O.conf_iters(5, 30000, 3, 1.0, 12345) == 12345
    d = synthetic.make_correspondences(3000, inlier_ratio=0.5, seed=9)
    a = O.ransac(d["src"], d["tgt"], conf=0.9995, max_iters=100000, round_size=1024, seed=3)
    assert a["iters_run"] == 1024  # exit is evaluated at round ends
    b = O.ransac(d["src"], d["tgt"], conf=1.0, max_iters=1024, round_size=256, seed=3)
    assert b["best_id"] == a["best_id"] and b["best_count"] == a["best_count"]
    O.set_threads(1)
    c = O.ransac(d["src"], d["tgt"], conf=1.0, max_iters=1024, round_size=1024, seed=3)
    O.set_threads(O.num_threads() if O.num_threads() > 1 else 8)
    assert c["best_id"] == a["best_id"] and np.array_equal(c["T"], a["T"])  # thread-count independent


def test_degenerate_inputs():
    src = np.zeros((2, 3), np.float32)
    r = O.ransac(src, src, m=3, max_iters=10)
    assert np.allclose(r["T"], np.eye(4)) and r["best_id"] == -1
    # no inliers anywhere: identity, like Open3D's fitness-0 result
    rng = np.random.default_rng(0)
    a = rng.uniform(-80, 80, (50, 3)).astype(np.float32)
    b = rng.uniform(500, 900, (50, 3)).astype(np.float32)
    r = O.ransac(a, b, m=3, use_elc=False, thr=1e-6, max_iters=50)
    assert np.allclose(r["T"], np.eye(4))


def test_refit_indexed_matches_direct():
    p = synthetic.make_pair(1200, seed=4, overlap=0.7)
    _, i1, _ = O.find_nn(p["feat0"], p["feat1"])
    i0 = np.arange(1200)
    T, k = O.refit_indexed(p["xyz0"], p["xyz1"], i0, i1, p["T_gt"], 0.6)
    assert k > 300
    assert np.allclose(T, p["T_gt"], atol=0.05)


@pytest.mark.parametrize("outlier", [0.5, 0.7])
def test_oracle_recovers_motion(outlier):
    d = synthetic.make_correspondences(4000, inlier_ratio=1 - outlier, seed=11)
    r = O.ransac(d["src"], d["tgt"], m=3, use_elc=True, thr=0.6, conf=1.0, max_iters=4096, round_size=4096, seed=51)
    from lidarregistration_b200 import metrics
    assert metrics.registration_success(r["T_refit"], d["T_gt"])
    assert 0 < r["n_passed"] < 4096


def test_refit_against_reference_weighted_procrustes(refit_golden):
    """a13: inliers of a coarse model at 0.6 m -> Kabsch, against the reference's own weighted_procrustes
    (DGR/util/procrustes.py:34-56, fp32 output) run on the same inputs with the mask as 0/1 weights."""
    for g in refit_golden:
        n = len(g["src"])
        idx = np.arange(n)
        cnt, mask = O.count_inliers(g["src"], g["tgt"], g["T_in"], 0.6, return_mask=True)
        assert np.array_equal(mask, g["mask"]) and cnt == int(g["mask"].sum())
        T, k = O.refit_indexed(g["src"], g["tgt"], idx, idx, g["T_in"], 0.6)
        assert k == cnt
        if k < 3:
            continue  # rank-deficient sets: both sides return *a* rotation, not a comparable one
        scale = 1.0 + np.abs(g["src"][g["mask"]]).max()
        assert np.abs(T[:3, :3] - g["R"]).max() < 5e-6, np.abs(T[:3, :3] - g["R"]).max()
        assert np.abs(T[:3, 3] - g["t"]).max() < 1e-5 * scale, np.abs(T[:3, 3] - g["t"]).max()


needs_ref = pytest.mark.skipif(not O.has_ref(), reason="oracle/_ref/libelc_ref.so not built (needs /root/reference)")


@needs_ref
def test_elc_against_compiled_reference_header():
    """The oracle's lro_elc vs the REFERENCE's own EdgeLenPreemptiveVerification::verifyModel
    (preemption_edge_length.h:71-128, compiled unmodified from the reference tree into oracle/_ref): random
    triplets / quadruplets on LiDAR-shaped correspondences (about half pass), duplicate indices, and edges
    constructed to sit exactly on / one ulp around the 0.9 boundary."""
    d = synthetic.make_correspondences(4000, inlier_ratio=0.5, seed=77)
    src, tgt = d["src"], d["tgt"]
    rng = np.random.default_rng(8)
    inl = np.flatnonzero(d["is_inlier"])
    n_pass = 0
    for m in (3, 4):
        for trial in range(6000):
            s = rng.choice(inl, m, replace=False) if trial % 2 else rng.integers(0, 4000, m)
            if trial % 50 == 0:
                s[1] = s[0]  # duplicates: 0 < 0 is false -> that edge passes
            ref = O.ref_elc(src, tgt, s)
            mine = O.elc(src[s].astype(np.float64), tgt[s].astype(np.float64), 0.9)
            assert ref == mine, (m, trial, s)
            n_pass += ref
    assert 2000 < n_pass < 10000  # both outcomes are exercised
    # boundary: target edge = source edge / 0.9 along one axis, nudged by single ulps of fp32 coordinates
    base = np.zeros((3, 3), np.float32)
    for k in range(400):
        L = np.float32(rng.uniform(1, 60))
        P = base.copy(); P[1, 0] = L; P[2, 1] = L
        Q = base.copy(); Q[1, 0] = np.float32(L / np.float32(0.9)); Q[2, 1] = L
        for nudge in (-2, -1, 0, 1, 2):
            Qn = Q.copy()
            for _ in range(abs(nudge)):
                Qn[1, 0] = np.nextafter(Qn[1, 0], np.float32(np.inf if nudge > 0 else -np.inf))
            assert O.ref_elc(P, Qn, [0, 1, 2]) == O.elc(P.astype(np.float64), Qn.astype(np.float64), 0.9)
