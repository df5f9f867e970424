"""Tensor-core inlier sweep (csrc/lr_score_tc.cuh): the proven error bound of the tcgen05 residual components
against fp64, exact counts through the banded fp64 recheck, and equality of the three sweep implementations.
The probe doubles as the micro-benchmark that establishes the accumulation-error constant (kAccKappa): its
measured worst ratio |d_tc - d_64| / E is written to gpurun_out/tc_probe_bound.json when that directory exists."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from lidarregistration_b200 import engine, synthetic  # noqa: E402
from oracle import lr_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_models(rng, H, yaw=180.0, rp=10.0, t=60.0):
    """[H,12] fp64 rows [R|t] of random rigid motions (Rz Ry Rx, wide ranges: the sweep sees wild hypotheses too)"""
    out = np.zeros((H, 12))
    for h in range(H):
        a, b, c = np.deg2rad(rng.uniform(-yaw, yaw)), np.deg2rad(rng.normal(0, rp)), np.deg2rad(rng.normal(0, rp))
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(c), -np.sin(c)], [0, np.sin(c), np.cos(c)]])
        T = np.zeros((3, 4))
        T[:, :3] = Rz @ Ry @ Rx
        T[:, 3] = rng.uniform(-t, t, 3)
        out[h] = T.reshape(-1)
    return out


def d64(src, tgt, models):
    """canonical fp64 residual components [H, n, 3]"""
    p, q = src.astype(np.float64), tgt.astype(np.float64)
    M = models.reshape(-1, 3, 4)
    return np.einsum("hab,nb->hna", M[:, :, :3], p) + M[:, None, :, 3] - q[None]


def probe_case(src, tgt, models):
    d, E, counts = engine.tc_probe(src, tgt, models, 0.6)
    n = src.shape[0]
    d = d.cpu().numpy()[:, :n, :]
    E = E.cpu().numpy()
    err = np.abs(d.astype(np.float64) - d64(src, tgt, models)).max(axis=(1, 2))
    return err, E, counts.cpu().numpy(), d


def test_probe_error_bound_and_exact_counts():
    rng = np.random.default_rng(7)
    report = {}
    worst = 0.0
    for name, n, H, noise, offset in [("lidar", 4000, 256, 0.1, 0.0), ("dense_near_threshold", 3000, 256, 0.45, 0.0),
                                      ("map_frame_10km", 3000, 128, 0.1, 10000.0), ("ragged", 1237, 130, 0.2, 0.0)]:
        d = synthetic.make_correspondences(n, inlier_ratio=0.5, seed=100 + n, noise=noise)
        src, tgt = d["src"].copy(), d["tgt"].copy()
        if offset:
            src += np.float32(offset)
            tgt -= np.float32(0.7 * offset)
        gt = np.asarray(d["T_gt"], dtype=np.float64)[:3, :].copy()
        if offset:  # the motion that maps the shifted source onto the shifted target
            gt[:, 3] = gt[:, 3] - 0.7 * offset - gt[:, :3] @ np.full(3, offset)
        models = random_models(rng, H)
        if offset:
            models[:, 3::4] += (-0.7 * offset - models.reshape(-1, 3, 4)[:, :, :3].sum(axis=2) * offset)
        # half of the models are small perturbations of the true motion: residuals crowd around the threshold
        for h in range(0, H, 2):
            T = gt.copy()
            T[:, 3] += rng.normal(0, 0.2, 3)
            models[h] = T.reshape(-1)
        err, E, counts, _ = probe_case(src, tgt, models)
        ratio = float((err / E).max())
        report[name] = {"max_err_m": float(err.max()), "min_bound_m": float(E.min()), "max_err_over_bound": ratio}
        worst = max(worst, ratio)
        assert ratio < 1.0, (name, ratio)
        ref = np.array([O.count_inliers(src, tgt, np.vstack([m.reshape(3, 4), [0, 0, 0, 1]]), 0.6) for m in models])
        assert np.array_equal(counts, ref), name
    report["worst_ratio"] = worst
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(report, open(os.path.join(out, "tc_probe_bound.json"), "w"), indent=1)
    # the accumulation term of the bound is meant to be generous: the worst observed error stays below half of it
    assert worst < 0.5, report


def test_probe_adversarial_pieces():
    """coordinates / model entries chosen so that every dropped piece product is as large as it can be
    (values just above a power of two plus half an fp16 ulp), and the largest in-range coordinates"""
    rng = np.random.default_rng(11)
    n, H = 2048, 128
    base = np.array([64.0, 32.0, 127.9, 100.03, 0.5, 3.999])
    src = (rng.choice(base, (n, 3)) * (1 + 2.0 ** -12 + 2.0 ** -23) * rng.choice([-1, 1], (n, 3))).astype(np.float32)
    tgt = (rng.choice(base, (n, 3)) * (1 + 2.0 ** -12 + 2.0 ** -23) * rng.choice([-1, 1], (n, 3))).astype(np.float32)
    models = random_models(rng, H, t=120.0)
    err, E, counts, _ = probe_case(src, tgt, models)
    assert (err / E).max() < 1.0
    # coordinates up to 3.2 km: relative to the operand frame (first correspondence, rounded to 1024 m) that is
    # |p~|_2 <= sqrt(3) (2 x 3.2 km + 512 m) = 12 km, still inside the fp16 range guard (15 km)
    src2 = (src * np.float32(25.0)).astype(np.float32)
    tgt2 = (tgt * np.float32(25.0)).astype(np.float32)
    err, E, counts, _ = probe_case(src2, tgt2, models)
    assert (err / E).max() < 1.0
    ref = np.array([O.count_inliers(src2, tgt2, np.vstack([m.reshape(3, 4), [0, 0, 0, 1]]), 0.6) for m in models])
    assert np.array_equal(counts, ref)


def test_out_of_range_falls_back_to_exact_fp64():
    """a cloud wider than the fp16 piece range (> 15 km from the operand frame): counts still exact"""
    d = synthetic.make_correspondences(600, inlier_ratio=0.5, seed=3)
    src, tgt = d["src"].copy(), d["tgt"].copy()
    src[5] += np.float32(40000.0)
    tgt[7] -= np.float32(50000.0)
    rng = np.random.default_rng(3)
    samples = rng.integers(0, 600, (500, 3)).astype(np.int32)
    engine.ransac_set_mode(0)
    counts, best, _ = engine.ransac_score_samples(src, tgt, samples, 0.6, True, 0.9)
    oc, ob = O.score_samples(src, tgt, samples, 0.6, True, 0.9)
    assert np.array_equal(counts.cpu().numpy(), oc) and best == ob


@pytest.mark.parametrize("m", [3, 4])
def test_three_sweeps_agree_with_oracle(m):
    """tensor-core sweep == fp32 sweep == fp32 sweep with early-out == oracle, counts bit for bit"""
    d = synthetic.make_correspondences(9000, inlier_ratio=0.3, seed=77)
    rng = np.random.default_rng(5)
    samples = rng.integers(0, 9000, (6000, m)).astype(np.int32)
    inl = np.flatnonzero(d["is_inlier"])
    samples[::3] = rng.choice(inl, (len(samples[::3]), m))  # plenty of all-inlier samples that pass ELC
    oc, ob = O.score_samples(d["src"], d["tgt"], samples, 0.6, True, 0.9)
    try:
        for mode in (0, 1, 2):
            engine.ransac_set_mode(mode)
            counts, best, _ = engine.ransac_score_samples(d["src"], d["tgt"], samples, 0.6, True, 0.9)
            assert np.array_equal(counts.cpu().numpy(), oc), mode
            assert best == ob, mode
    finally:
        engine.ransac_set_mode(0)


def test_full_loop_modes_identical():
    d = synthetic.make_correspondences(20000, inlier_ratio=0.25, seed=12)
    outs = []
    try:
        for mode in (0, 1, 2):
            engine.ransac_set_mode(mode)
            for conf, iters in ((1.0, 200000), (0.9995, 400000)):
                p = engine.make_params(confidence=conf, max_iters=iters, seed=9, use_elc=True, round_size=32768)
                r = engine.ransac_rigid(d["src"], d["tgt"], p, want_mask=True)
                outs.append((mode, conf, r))
    finally:
        engine.ransac_set_mode(0)
    for conf in (1.0, 0.9995):
        rs = [r for (_, c, r) in outs if c == conf]
        for r in rs[1:]:
            for k in ("best_id", "best_count", "iters_run", "n_scored", "refit_count"):
                assert r[k] == rs[0][k], (conf, k)
            assert np.array_equal(r["T"], rs[0]["T"]) and np.array_equal(r["T_refit"], rs[0]["T_refit"])
            assert torch.equal(r["mask"], rs[0]["mask"])
    assert outs[0][2]["n_rechecked"] >= 0


def test_refit_is_bit_reproducible():
    """k_finish adds the per-block sums in a fixed order: T_refit has no run-to-run last-bit noise"""
    d = synthetic.make_correspondences(30000, inlier_ratio=0.3, seed=4)
    p = engine.make_params(confidence=1.0, max_iters=100000, seed=2, use_elc=True)
    ref = engine.ransac_rigid(d["src"], d["tgt"], p)
    for _ in range(4):
        r = engine.ransac_rigid(d["src"], d["tgt"], p)
        assert np.array_equal(r["T_refit"], ref["T_refit"]) and r["refit_count"] == ref["refit_count"]
    o = O.ransac(d["src"], d["tgt"], m=3, sampler=0, use_elc=True, thr=0.6, conf=1.0, max_iters=100000, round_size=65536,
                 seed=2, refit=True)
    assert ref["best_id"] == o["best_id"] and ref["best_count"] == o["best_count"]
    assert np.abs(ref["T_refit"][:3, :3] - o["T_refit"][:3, :3]).max() < 1e-5
    assert np.abs(ref["T_refit"][:3, 3] - o["T_refit"][:3, 3]).max() < 1e-4
