"""GC-RANSAC semantics on the GPU (MSAC selection, local optimisation, iterated least squares)
vs the oracle's lro_ransac_gc, through the C ABI.  Integer results (selected hypothesis, every q, inlier numbers,
improvement counters of the exact stages) are bit-exact; models of the exact stages are identical doubles; the
least-squares stage (block-reduced sums) is compared at the north star's tolerance."""
import sys

import numpy as np
import pytest
import torch

from lidarregistration_b200 import engine, metrics, synthetic
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu
ROT_TOL, TRANS_TOL = 1e-5, 1e-4  # BASELINE.json north_star
THR = 0.6


def close_T(a, b, rt=ROT_TOL, tt=TRANS_TOL):
    return np.abs(a[:3, :3] - b[:3, :3]).max() < rt and np.abs(a[:3, 3] - b[:3, 3]).max() < tt


@pytest.mark.parametrize("m", [3, 4])
@pytest.mark.parametrize("use_elc", [True, False])
def test_fed_samples_msac_bit_exact(m, use_elc):
    d = synthetic.make_correspondences(6000, inlier_ratio=0.3, seed=21)
    rng = np.random.default_rng(m)
    H = 20000
    samples = rng.integers(0, 6000, (H, m)).astype(np.int32)
    samples[:50, 1] = samples[:50, 0]
    samples[50:60] = samples[50:60, :1]
    inl_idx = np.flatnonzero(d["is_inlier"])
    samples[7000, :3] = inl_idx[[1, 300, 900]]
    samples[9000] = samples[7000]  # a tie between two good hypotheses: the lower index wins
    scores, inl, best = engine.ransac_score_samples_msac(d["src"], d["tgt"], samples, THR, use_elc, 0.9)
    os_, oi, ob = O.score_samples_msac(d["src"], d["tgt"], samples, THR, use_elc, 0.9)
    assert np.array_equal(scores.cpu().numpy(), os_)
    assert np.array_equal(inl.cpu().numpy(), oi)
    assert best == ob


def test_fed_samples_msac_many_survivors_and_point_splits():
    """ELC off on a small pair -> every sample survives (slot blocks >> CTAs, no point split); ELC on with few
    samples -> few survivors (points split over CTAs, integer atomics): same q either way."""
    d = synthetic.make_correspondences(30000, inlier_ratio=0.3, seed=51 + 3000)
    rng = np.random.default_rng(0)
    for H, elc in ((300, False), (3000, True), (150000, True)):
        samples = rng.integers(0, 30000, (H, 3)).astype(np.int32)
        scores, inl, best = engine.ransac_score_samples_msac(d["src"], d["tgt"], samples, THR, elc, 0.9)
        scores, inl = scores.cpu().numpy(), inl.cpu().numpy()
        sub = np.unique(np.concatenate([np.arange(0, H, max(1, H // 400)), [max(best, 0)]]))
        os_, oi, _ = O.score_samples_msac(d["src"], d["tgt"], samples[sub], THR, elc, 0.9)
        assert np.array_equal(scores[sub], os_) and np.array_equal(inl[sub], oi)
        assert best == (int(np.argmax(scores)) if scores.max() > 0 else -1)


@pytest.mark.parametrize("sampler", [engine.SAMPLER_UNIFORM, engine.SAMPLER_PROSAC])
def test_gc_stages_match_oracle(sampler):
    d = synthetic.make_correspondences(8000, inlier_ratio=0.25, seed=77)
    kw = dict(threshold=THR, confidence=1.0, max_iters=40000, seed=5, sampler=sampler, use_elc=True,
              scoring=engine.SCORE_MSAC, lo_trials=20)
    okw = dict(m=3, sampler=sampler, use_elc=True, thr=THR, conf=1.0, max_iters=40000, seed=5, lo_trials=20)
    # 1. selection only
    g = engine.ransac_rigid(d["src"], d["tgt"], engine.make_params(lo_rounds=0, lsq_iters=0, **kw), want_mask=True)
    o = O.ransac_gc(d["src"], d["tgt"], lo_rounds=0, lsq_iters=0, return_mask=True, **okw)
    assert (g["best_id"], g["best_score"], g["best_count"]) == (o["best_id"], o["best_score"], o["best_inliers"])
    assert g["iters_run"] == o["iters_run"] and g["n_scored"] == o["n_passed"]
    assert g["final_score"] == g["lo_score"] == g["best_score"]
    assert np.array_equal(g["T"], o["T"])
    assert np.array_equal(g["mask"].cpu().numpy(), o["mask"]) and g["refit_count"] == o["refit_count"]
    assert close_T(g["T_refit"], o["T_refit"])
    # 2. + local optimisation: still exact (small-sample Kabsch in draw order, integer scores)
    g = engine.ransac_rigid(d["src"], d["tgt"], engine.make_params(lo_rounds=10, lsq_iters=0, **kw))
    o = O.ransac_gc(d["src"], d["tgt"], lo_rounds=10, lsq_iters=0, **okw)
    assert o["lo_improved"] >= 1  # the case exercises the stage
    assert (g["lo_score"], g["lo_improved"], g["final_score"]) == (o["lo_score"], o["lo_improved"], o["final_score"])
    assert np.array_equal(g["T"], o["T"])
    # 3. + iterated least squares: sums over ~2000 inliers are block-reduced on the GPU
    g = engine.ransac_rigid(d["src"], d["tgt"], engine.make_params(lo_rounds=10, lsq_iters=10, **kw))
    o = O.ransac_gc(d["src"], d["tgt"], lo_rounds=10, lsq_iters=10, **okw)
    assert g["lo_score"] == o["lo_score"] and g["lsq_improved"] == o["lsq_improved"]
    assert abs(g["final_score"] - o["final_score"]) <= 1e-7 * o["final_score"]
    assert close_T(g["T"], o["T"], 1e-9, 1e-8)
    assert metrics.registration_success(g["T"], d["T_gt"])
    # the reported score is the returned model's
    assert O.msac_q(d["src"], d["tgt"], g["T"], THR)[0] == g["final_score"]


def test_gc_confidence_exit_and_rounds():
    d = synthetic.make_correspondences(5000, inlier_ratio=0.5, seed=31)
    p = engine.make_params(threshold=THR, confidence=0.999, max_iters=500000, seed=9, use_elc=True, round_size=1024,
                           scoring=engine.SCORE_MSAC, lo_rounds=3, lo_trials=8, lsq_iters=2)
    g = engine.ransac_rigid(d["src"], d["tgt"], p)
    o = O.ransac_gc(d["src"], d["tgt"], thr=THR, conf=0.999, max_iters=500000, seed=9, round_size=1024, lo_rounds=3,
                    lo_trials=8, lsq_iters=2)
    assert g["iters_run"] == o["iters_run"] < 500000
    assert (g["best_id"], g["best_score"], g["best_count"]) == (o["best_id"], o["best_score"], o["best_inliers"])
    assert (g["lo_score"], g["lo_improved"]) == (o["lo_score"], o["lo_improved"])
    assert close_T(g["T"], o["T"], 1e-9, 1e-8)


def test_gc_degenerate_inputs():
    p = engine.make_params(threshold=THR, max_iters=1000, scoring=engine.SCORE_MSAC, lo_rounds=2, lsq_iters=2)
    rng = np.random.default_rng(0)
    # fewer correspondences than the sample size -> identity (same rule as count scoring)
    r = engine.ransac_rigid(rng.normal(size=(2, 3)).astype(np.float32), rng.normal(size=(2, 3)).astype(np.float32), p)
    assert np.array_equal(r["T"], np.eye(4)) and r["best_id"] == -1
    # nothing can score: unrelated clouds at a tiny threshold
    src = rng.uniform(-50, 50, (500, 3)).astype(np.float32)
    tgt = (rng.uniform(-50, 50, (500, 3)) + 1e4 * np.arange(500)[:, None]).astype(np.float32)
    q = engine.make_params(threshold=1e-6, max_iters=500, use_elc=False, scoring=engine.SCORE_MSAC, lo_rounds=2,
                           lsq_iters=2)
    g = engine.ransac_rigid(src, tgt, q, want_mask=True)
    o = O.ransac_gc(src, tgt, thr=1e-6, max_iters=500, use_elc=False, lo_rounds=2, lsq_iters=2)
    assert (g["best_id"], g["best_score"], g["final_score"]) == (o["best_id"], o["best_score"], o["final_score"])
    assert close_T(g["T"], o["T"], 1e-9, 1e-8)
    # exactly three correspondences: LO has nothing to draw from (|L| <= m), the model is the sample's
    d = synthetic.make_correspondences(3, inlier_ratio=1.0, seed=1, noise=0.0)
    g = engine.ransac_rigid(d["src"], d["tgt"], engine.make_params(threshold=THR, max_iters=10, use_elc=False,
                                                                   scoring=engine.SCORE_MSAC, lo_rounds=5,
                                                                   lsq_iters=0))
    o = O.ransac_gc(d["src"], d["tgt"], thr=THR, max_iters=10, use_elc=False, lo_rounds=5, lsq_iters=0)
    assert g["lo_improved"] == o["lo_improved"] == 0 and np.array_equal(g["T"], o["T"])
    # sharding packs (count, id): MSAC runs are refused loudly, not silently scored by count
    key = torch.zeros(1, dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="LR_SCORE_COUNT"):
        engine.ransac_shard(engine.to_dev_f32(src), engine.to_dev_f32(tgt), q, 0, 100, key)


def test_gc_batch_entry_equals_single_calls():
    p = engine.make_params(threshold=THR, max_iters=30000, seed=3, scoring=engine.SCORE_MSAC, lo_rounds=4,
                           lo_trials=12, lsq_iters=3)
    pairs = []
    for k, n in enumerate((5000, 2, 7000, 3100, 6000)):
        d = synthetic.make_correspondences(n, inlier_ratio=0.3, seed=200 + k)
        pairs.append((engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])))
    batch = engine.ransac_rigid_batch(pairs, p)
    for (a, b), r in zip(pairs, batch):
        s = engine.ransac_rigid(a, b, p)
        for key in ("best_id", "best_score", "best_count", "lo_score", "lo_improved", "iters_run", "n_scored"):
            assert r[key] == s[key], key
        assert close_T(r["T"], s["T"], 1e-9, 1e-8)
    # count-scoring runs after MSAC runs in the same arenas are unaffected
    pc = engine.make_params(threshold=THR, max_iters=30000, seed=3)
    a, b = pairs[0]
    g = engine.ransac_rigid(a, b, pc)
    o = O.ransac(a.cpu().numpy(), b.cpu().numpy(), thr=THR, max_iters=30000, seed=3)
    assert (g["best_id"], g["best_count"]) == (o["best_id"], o["best_count"]) and g["best_score"] == 0


def test_reference_interface_with_msac_scoring():
    """findRigidTransform / FR with GC_scoring = MSAC: pygcransac's return convention, GC_LO switch, accuracy."""
    from lidarregistration_b200.algorithms import FR
    G = sys.modules["lidarregistration_b200.algorithms.GC_RANSAC"]  # the package re-exports the function under this name
    d = synthetic.make_correspondences(6000, inlier_ratio=0.3, seed=41)
    common = dict(threshold=THR, conf=1.0, spatial_coherence_weight=0.0, max_iters=30000, use_sprt=True,
                  min_inlier_ratio_for_sprt=-1, sampler=0, neighborhood_size=20, seed=7)
    pose, mask = G.findRigidTransform(d["src"], d["tgt"], neighborhood=0, scoring="MSAC", **common)
    o = O.ransac_gc(d["src"], d["tgt"], thr=THR, max_iters=30000, seed=7, return_mask=True)
    assert close_T(pose.T, o["T"], 1e-9, 1e-8) and np.array_equal(mask, o["mask"])
    pose_nolo, _ = G.findRigidTransform(d["src"], d["tgt"], neighborhood=1, scoring="MSAC", **common)
    # --GC_LO False only switches the graph-cut rounds off (gcransac_python.cpp:518-521); the finishing iterated
    # least squares still runs (SURVEY App. A "Finish")
    o2 = O.ransac_gc(d["src"], d["tgt"], thr=THR, max_iters=30000, seed=7, lo_rounds=0)
    assert close_T(pose_nolo.T, o2["T"], 1e-9, 1e-8)
    with pytest.raises(NotImplementedError):
        G.findRigidTransform(d["src"], d["tgt"], neighborhood=0, scoring="MSAC",
                             **dict(common, spatial_coherence_weight=0.1))
    # through FR() on a full synthetic pair
    p = synthetic.make_pair(4000, seed=61, overlap=0.7)

    class A:
        pass

    args = A()
    for k, v in dict(mode="MNN", iters=30000, codebase="GC", prosac=True, spatial_coherence_weight=0.0, GC_conf=0.9995,
                     fast_rejection="ELC", GC_LO=True, GPF_factor=2.0, GPF_grid_wid=10, GPF_max_matches=10 ** 9,
                     seed=51, GC_scoring="MSAC").items():
        setattr(args, k, v)
    t = [torch.from_numpy(p[k]) for k in ("xyz0", "xyz1", "feat0", "feat1")]
    T = FR(*t, args, p["T_gt"])[0]
    assert metrics.registration_success(T, p["T_gt"])
    assert metrics.translation_error_cm(T, p["T_gt"]) < 10.0


def test_zy_refit_against_reference_weighted_procrustes(refit_golden):
    """a13 on the GPU (lr_refit_indexed) against the reference's own weighted_procrustes run on the same inputs
    (tests/golden/refit_ref.npz; the oracle agrees with it to 1e-7 / 3e-6 m, tests/test_oracle_ransac.py)."""
    for g in refit_golden:
        n = len(g["src"])
        idx = np.arange(n)
        T, k = engine.refit_indexed(g["src"], g["tgt"], idx, idx, g["T_in"], 0.6)
        assert k == int(g["mask"].sum())
        if k < 3:
            continue
        assert np.abs(T[:3, :3] - g["R"]).max() < ROT_TOL and np.abs(T[:3, 3] - g["t"]).max() < TRANS_TOL  # north star

@pytest.mark.skipif(not O.has_ref(), reason="oracle/_ref/libelc_ref.so not built (needs /root/reference)")
def test_zz_cuda_elc_against_compiled_reference_header():
    """k_gen's edge-length decision (elc_pass_fast: squared lengths, sqrt form only near equality) against the
    REFERENCE's own verifyModel (preemption_edge_length.h:71-128 compiled into oracle/_ref), through the fed-sample
    hook: count == -1 <=> the reference rejects.  Random triplets plus edges placed on / within two fp32 ulps of
    the 0.9 boundary.  (Last test of the suite on purpose.)"""
    d = synthetic.make_correspondences(4000, inlier_ratio=0.5, seed=77)
    rng = np.random.default_rng(9)
    inl = np.flatnonzero(d["is_inlier"])
    P_extra, Q_extra = [], []
    for k in range(300):
        L = np.float32(rng.uniform(1, 60))
        for nudge in (-2, -1, 0, 1, 2):
            P = np.zeros((3, 3), np.float32); P[1, 0] = L; P[2, 1] = L
            Q = np.zeros((3, 3), np.float32); Q[1, 0] = np.float32(L / np.float32(0.9)); Q[2, 1] = L
            for _ in range(abs(nudge)):
                Q[1, 0] = np.nextafter(Q[1, 0], np.float32(np.inf if nudge > 0 else -np.inf))
            P_extra.append(P), Q_extra.append(Q)
    src = np.concatenate([d["src"]] + P_extra).astype(np.float32)
    tgt = np.concatenate([d["tgt"]] + Q_extra).astype(np.float32)
    rand = np.stack([rng.choice(inl, 3, replace=False) if t % 2 else rng.integers(0, 4000, 3) for t in range(6000)])
    edge = 4000 + 3 * np.arange(len(P_extra))[:, None] + np.arange(3)[None, :]
    samples = np.concatenate([rand, edge]).astype(np.int32)
    counts, _, _ = engine.ransac_score_samples(src, tgt, samples, THR, True, 0.9)
    rejected = counts.cpu().numpy() < 0
    ref = np.array([O.ref_elc(src, tgt, s) for s in samples])
    assert np.array_equal(rejected, ~ref)
    assert 0.2 < ref[:6000].mean() < 0.8 and 0 < ref[6000:].sum() < len(edge)  # both outcomes, also at the boundary
