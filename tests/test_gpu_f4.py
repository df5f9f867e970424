"""SURVEY 8(f4) on the GPU: the device-resident ICP refinement (Experiments/test.py:183-188) and PointDSC's seed
scoring (Experiments/models/PointDSC.py:293-336) against the oracle and the reference-generated fixture
tests/golden/seeds_ref.npz.  Integer results (nearest-neighbour indices, counts, selected seed, labels) are exact;
the per-seed weighted Kabsch is bit-identical to the oracle's (same operations in the same order)."""
import time

import numpy as np
import pytest
import torch

from lidarregistration_b200 import engine, synthetic
from lidarregistration_b200.algorithms import (registration_icp, registration_icp_bruteforce, score_seeds,
                                               seedwise_transforms)
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu


def small_T(ang_deg=1.0, t=(0.2, -0.1, 0.05)):
    a = np.deg2rad(ang_deg)
    T = np.eye(4)
    T[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    T[:3, 3] = t
    return T


@pytest.mark.parametrize("n,m,offset", [(5000, 6000, 0.0), (1, 1, 0.0), (777, 3, 0.0), (4000, 4000, 250000.0)])
def test_nn3d_radius_exact(n, m, offset):
    """hashed-grid search == brute force: same index (ties -> lowest), same fp64 squared distance, -1 outside the radius"""
    rng = np.random.default_rng(n + m)
    tgt = rng.uniform(-30, 30, (m, 3)).astype(np.float64)
    tgt[:, 2] *= 0.1
    src = tgt[rng.integers(0, m, n)] + rng.normal(0, 0.25, (n, 3))
    src[: n // 10] += 500.0                       # nothing within the radius
    if m > 200:
        tgt[100] = tgt[7]                          # duplicate target points
        tgt[150] = tgt[7]
        src[n // 2] = tgt[7] + 0.01
    src, tgt = (src + offset).astype(np.float32), (tgt + offset).astype(np.float32)
    T = small_T(0.02 if offset else 1.0)
    if offset:  # keep the moved points near the targets although the frame is far from the origin
        c = np.array([offset, offset, offset])
        T[:3, 3] += c - T[:3, :3] @ c
    idx, d2 = engine.nn3d_radius(src, tgt, T, 0.6)
    oi, od = O.nn3d_radius(src, tgt, T, 0.6)
    assert np.array_equal(idx.cpu().numpy(), oi)
    assert np.array_equal(d2.cpu().numpy(), od)
    assert (oi >= 0).sum() > 0 or n < 10


def test_nn3d_radius_empty_target():
    src = np.zeros((10, 3), np.float32)
    idx, _ = engine.nn3d_radius(src, np.zeros((0, 3), np.float32), np.eye(4), 0.6)
    assert (idx.cpu().numpy() == -1).all()


def test_icp_refine_matches_oracle():
    p = synthetic.make_pair(8000, seed=4242, overlap=0.7)
    T0 = small_T(1.5, (0.25, -0.2, 0.1)) @ p["T_gt"]
    res = registration_icp(p["xyz0"], p["xyz1"], 0.6, T0)
    To, fo, ro, ito = O.icp(p["xyz0"], p["xyz1"], 0.6, T0)
    assert res.iterations == ito and res.fitness == fo  # integer count / n
    assert abs(res.inlier_rmse - ro) < 1e-9 and np.abs(res.transformation - To).max() < 1e-8
    # the evaluation of the initial transform alone
    r0 = registration_icp(p["xyz0"], p["xyz1"], 0.6, T0, max_iteration=0)
    T1, f1, e1, it1 = O.icp(p["xyz0"], p["xyz1"], 0.6, T0, max_iteration=0)
    assert r0.iterations == 0 == it1 and r0.fitness == f1 and abs(r0.inlier_rmse - e1) < 1e-9
    assert np.array_equal(r0.transformation, T0)
    # a capped run stops at the cap
    r3 = registration_icp(p["xyz0"], p["xyz1"], 0.6, T0, max_iteration=3)
    T3, f3, e3, it3 = O.icp(p["xyz0"], p["xyz1"], 0.6, T0, max_iteration=3)
    assert r3.iterations == it3 == 3 and r3.fitness == f3 and np.abs(r3.transformation - T3).max() < 1e-8
    # bit-reproducible run to run (fixed-order reductions)
    again = registration_icp(p["xyz0"], p["xyz1"], 0.6, T0)
    assert np.array_equal(again.transformation, res.transformation) and again.inlier_rmse == res.inlier_rmse
    # same fixed point as the brute-force composition of round 2 (different nearest-neighbour arithmetic: tolerance)
    bf = registration_icp_bruteforce(p["xyz0"], p["xyz1"], 0.6, T0)
    assert np.abs(bf.transformation - res.transformation).max() < 1e-3 and abs(bf.fitness - res.fitness) < 2e-3


def test_icp_refine_degenerate_inputs():
    z = np.zeros((0, 3), np.float32)
    pts = np.random.default_rng(1).uniform(-5, 5, (50, 3)).astype(np.float32)
    T0 = small_T()
    for a, b in ((z, pts), (pts, z)):
        r = registration_icp(a, b, 0.6, T0)
        assert np.array_equal(r.transformation, T0) and r.fitness == 0.0 and r.iterations == 0
    far = registration_icp(pts, pts + 100.0, 0.6, None)  # no pair inside the distance: Kabsch of nothing = identity
    assert far.fitness == 0.0 and np.array_equal(far.transformation[:3, :3], np.eye(3))


def test_icp_speed_vs_bruteforce():
    p = synthetic.make_pair(25000, seed=99, overlap=0.8)
    T0 = small_T(1.0, (0.2, -0.15, 0.05)) @ p["T_gt"]
    a, b = engine.to_dev_f32(p["xyz0"]), engine.to_dev_f32(p["xyz1"])
    out = {}
    for name, fn in (("grid", registration_icp), ("bruteforce", registration_icp_bruteforce)):
        fn(a, b, 0.6, T0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn(a, b, 0.6, T0)
        torch.cuda.synchronize()
        out[name] = ((time.perf_counter() - t0) * 1e3, r.iterations)
    print("ICP 25k points: grid %.2f ms (%d it), brute force %.2f ms (%d it)" % (out["grid"] + out["bruteforce"]))
    assert out["grid"][0] < out["bruteforce"][0]


def test_weighted_kabsch_batch_bit_exact(seeds_golden):
    for g in seeds_golden:
        T = engine.kabsch_weighted_batch(g["A"], g["B"], g["w"]).cpu().numpy()
        for s in range(0, len(T), 7):
            assert np.array_equal(T[s], O.kabsch_weighted(g["A"][s], g["B"][s], g["w"][s])), s
        T1 = engine.kabsch_weighted_batch(g["A"][:5], g["B"][:5], None).cpu().numpy()
        assert np.array_equal(T1[3], O.kabsch_weighted(g["A"][3], g["B"][3], None))
        # the reference's fp32 result (models/common.py:7-45) within its own precision
        ok = [s for s in range(len(T)) if np.allclose(T[s][:3, :3], g["trans"][s][:3, :3], atol=5e-3)]
        assert len(ok) >= 0.95 * len(T)
        assert np.allclose(seedwise_transforms(g["A"], g["B"], g["w"]).cpu().numpy(), T.astype(np.float32))


def test_seed_scoring_exact_and_against_reference(seeds_golden):
    for g in seeds_golden:
        src, tgt, thr = g["src"], g["tgt"], g["threshold"]
        trans = g["trans"].astype(np.float64)
        res = engine.seeds_score(src, tgt, trans, thr, want_refit=True)
        counts, best, labels = O.seeds_score(src, tgt, trans, thr, return_labels=True)
        assert np.array_equal(res["counts"].cpu().numpy(), counts)       # the tensor sweep's counts are exact
        assert res["best"] == best and res["best_count"] == counts[best]
        assert np.array_equal(res["labels"].cpu().numpy(), labels)
        assert np.array_equal(res["T"], trans[best])
        Tr, k = O.refit_indexed(src, tgt, np.arange(len(src)), np.arange(len(src)), trans[best], thr)
        assert k == counts[best] and np.abs(res["T_refit"] - Tr).max() < 1e-9
        # the reference's own outputs (fp32): identical apart from residuals within fp32 rounding of the threshold
        n = len(src)
        ref_counts = np.rint(g["fitness"].astype(np.float64) * n).astype(np.int64)
        assert np.abs(res["counts"].cpu().numpy() - ref_counts).max() <= 3
        fit, final_trans, final_labels, b = score_seeds(torch.from_numpy(g["trans"])[None], torch.from_numpy(src)[None],
                                                        torch.from_numpy(tgt)[None], thr)
        assert fit.shape == (1, len(trans)) and final_trans.shape == (1, 4, 4) and final_labels.shape == (1, n)
        assert np.allclose(fit[0].numpy(), g["fitness"], atol=3.5 / n)
        if b == int(np.argmax(g["fitness"])):
            assert np.allclose(final_trans[0].numpy(), g["final_trans"])
            assert (final_labels[0].numpy() != g["final_labels"]).sum() <= 3


def test_seed_scoring_zero_inlier_winner_and_many_seeds():
    """all seeds wrong: the arg-max of all-zero fitness is seed 0, whose transform is returned (PointDSC.py:326-329);
    and a seed set larger than one 128-row block with ties"""
    d = synthetic.make_correspondences(1500, 0.3, seed=5)
    far = np.tile(np.eye(4), (5, 1, 1))
    far[:, :3, 3] = 1e4
    r = engine.seeds_score(d["src"], d["tgt"], far, 0.6)
    assert r["best"] == 0 and r["best_count"] == 0 and np.array_equal(r["T"], far[0]) and not r["labels"].any()
    rng = np.random.default_rng(2)
    models = np.tile(d["T_gt"], (700, 1, 1))
    models[:, :3, 3] += rng.normal(0, 0.3, (700, 3))
    models[400] = d["T_gt"]
    models[123] = d["T_gt"]   # two identical best seeds: the first wins
    r = engine.seeds_score(d["src"], d["tgt"], models, 0.6)
    counts, best = O.seeds_score(d["src"], d["tgt"], models, 0.6)
    assert np.array_equal(r["counts"].cpu().numpy(), counts) and r["best"] == best
    assert counts[123] == counts[400] and (best <= 123)
