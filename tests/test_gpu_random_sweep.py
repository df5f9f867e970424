"""Randomised parity sweep: many small random configurations of both halves of the path, GPU vs oracle."""
import numpy as np
import pytest
import torch

from lidarregistration_b200 import engine, synthetic
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu


def test_matching_random_shapes_and_degeneracies():
    rng = np.random.default_rng(2024)
    for trial in range(40):
        N, M = int(rng.integers(1, 900)), int(rng.integers(2, 900))
        scale = float(10.0 ** rng.uniform(-3, 3))
        f0 = (rng.standard_normal((N, 32)) * scale).astype(np.float32)
        f1 = (rng.standard_normal((M, 32)) * scale).astype(np.float32)
        kind = trial % 5
        if kind == 1:  # many duplicates
            f1[rng.integers(0, M, M // 2)] = f1[0]
            f0[rng.integers(0, N, max(N // 3, 1))] = f1[0]
        elif kind == 2:  # quantised features: lots of exact ties
            f0, f1 = np.round(f0 / scale * 2) * scale / 2, np.round(f1 / scale * 2) * scale / 2
        elif kind == 3:  # one zero row, mixed norms
            f1[M // 2] = 0
            f0 *= rng.uniform(0.01, 100, (N, 1)).astype(np.float32)
        elif kind == 4:  # nearly identical sets
            f1[:min(N, M)] = f0[:min(N, M)] + (1e-6 * scale * rng.standard_normal((min(N, M), 32))).astype(np.float32)
        f0, f1 = np.ascontiguousarray(f0, np.float32), np.ascontiguousarray(f1, np.float32)
        i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
        _, o1, o2 = O.find_nn(f0, f1, return_2nd=True)
        assert np.array_equal(i1.cpu().numpy(), o1), (trial, N, M, kind)
        assert np.array_equal(i2.cpu().numpy(), o2), (trial, N, M, kind)
        mi, mj = engine.match_mutual(f0, f1, i1)
        oi, oj = O.nn_to_mutual(f0, f1, o1)
        assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(mj.cpu().numpy(), oj), (trial, N, M, kind)


def test_ransac_random_configurations():
    rng = np.random.default_rng(7)
    for trial in range(30):
        n = int(rng.integers(4, 3000))
        m = int(rng.choice([3, 4]))
        sampler = int(rng.choice([0, 1, 2]))
        use_elc = bool(rng.integers(0, 2))
        conf = float(rng.choice([1.0, 0.999, 0.9]))
        R = int(rng.choice([256, 1000, 4096]))
        iters = int(rng.integers(1, 12000))
        thr = float(rng.choice([0.3, 0.6, 2.0]))
        seed = int(rng.integers(0, 2 ** 31))
        d = synthetic.make_correspondences(n, inlier_ratio=float(rng.uniform(0.0, 0.9)), seed=seed % 100000)
        p = engine.make_params(threshold=thr, confidence=conf, max_iters=iters, seed=seed, sample_size=m, sampler=sampler,
                               use_elc=use_elc, round_size=R)
        r = engine.ransac_rigid(d["src"], d["tgt"], p, want_mask=True)
        ref = O.ransac(d["src"], d["tgt"], m=m, sampler=sampler, use_elc=use_elc, thr=thr, conf=conf, max_iters=iters,
                       round_size=R, seed=seed, return_mask=True)
        key = (trial, n, m, sampler, use_elc, conf, R, iters, thr)
        assert r["iters_run"] == ref["iters_run"], key
        assert r["best_count"] == ref["best_count"] and r["best_id"] == ref["best_id"], key
        assert r["n_scored"] == ref["n_passed"], key
        assert np.array_equal(r["T"], ref["T"]), key
        assert np.array_equal(r["mask"].cpu().numpy(), ref["mask"]), key
        assert np.abs(r["T_refit"] - ref["T_refit"]).max() < 1e-4, key
