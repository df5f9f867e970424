"""CUDA RANSAC vs the oracle, through the C ABI: fed samples (bit-exact counts / selection /
models), the sampler, the full loop, confidence exit, refit, degenerate inputs."""
import numpy as np
import pytest
import torch

from lidarregistration_b200 import engine, metrics, synthetic
from oracle import lr_oracle as O

pytestmark = pytest.mark.gpu
ROT_TOL, TRANS_TOL = 1e-5, 1e-4  # BASELINE.json north_star


def close_T(a, b):
    return np.abs(a[:3, :3] - b[:3, :3]).max() < ROT_TOL and np.abs(a[:3, 3] - b[:3, 3]).max() < TRANS_TOL


@pytest.mark.parametrize("m", [3, 4])
@pytest.mark.parametrize("use_elc", [True, False])
def test_fed_samples_bit_exact(m, use_elc):
    d = synthetic.make_correspondences(6000, inlier_ratio=0.3, seed=21)
    rng = np.random.default_rng(m)
    H = 20000
    samples = rng.integers(0, 6000, (H, m)).astype(np.int32)
    samples[:50, 1] = samples[:50, 0]  # repeated indices (Open3D draws with replacement)
    samples[50:60] = samples[50:60, :1]  # fully degenerate
    counts, best, models = engine.ransac_score_samples(d["src"], d["tgt"], samples, 0.6, use_elc, 0.9,
                                                       want_models=True)
    oc, ob, om = O.score_samples(d["src"], d["tgt"], samples, 0.6, use_elc, 0.9, return_models=True)
    assert np.array_equal(counts.cpu().numpy(), oc)
    assert best == ob
    assert np.array_equal(models.cpu().numpy(), om)  # same fp64 operation order -> identical models


def test_fed_triplets_cfg3_size():
    """north_star minimum slice: 100k fed triplets on a 30k-correspondence pair."""
    d = synthetic.make_correspondences(30000, inlier_ratio=0.3, seed=51 + 3000)
    rng = np.random.default_rng(0)
    H = 100000
    samples = rng.integers(0, 30000, (H, 3)).astype(np.int32)
    counts, best, _ = engine.ransac_score_samples(d["src"], d["tgt"], samples, 0.6, True, 0.9)
    counts = counts.cpu().numpy()
    sub = np.concatenate([np.arange(0, H, 97), [best]])
    oc, _ = O.score_samples(d["src"], d["tgt"], samples[sub], 0.6, True, 0.9)
    assert np.array_equal(counts[sub], oc)
    assert best == int(np.argmax(counts)) and counts[best] > 0.25 * 30000
    passed = counts >= 0
    assert 0.005 < passed.mean() < 0.2  # ELC pass rate at 70 % outliers (~3 %)


@pytest.mark.parametrize("sampler,m", [(engine.SAMPLER_UNIFORM, 3), (engine.SAMPLER_UNIFORM, 4),
                                       (engine.SAMPLER_REPLACE, 4), (engine.SAMPLER_REPLACE, 3),
                                       (engine.SAMPLER_PROSAC, 3), (engine.SAMPLER_PROSAC, 4)])
def test_sampler_matches_oracle(sampler, m):
    p = engine.make_params(sample_size=m, sampler=sampler, seed=1234)
    for n in (m, 37, 30000):
        got = engine.ransac_sample(p, n, 1000, 4096).cpu().numpy()
        if sampler == engine.SAMPLER_PROSAC and n > m:  # also the hand-over to uniform after T_N = 100000 draws
            late = engine.ransac_sample(p, n, 99990, 32).cpu().numpy()
            g = O.prosac_growth(n, m)
            assert np.array_equal(late, np.stack([O.sample(1234, 99990 + h, sampler, m, n, g) for h in range(32)]))
        growth = O.prosac_growth(n, m) if sampler == engine.SAMPLER_PROSAC else None
        want = np.stack([O.sample(1234, 1000 + h, sampler, m, n, growth) for h in range(0, 4096, 16)])
        assert np.array_equal(got[::16], want)


@pytest.mark.parametrize("inlier,use_elc,m,sampler", [(0.3, True, 3, 0), (0.3, False, 3, 0), (0.5, True, 4, 2),
                                                      (0.3, True, 3, 1)])
def test_full_loop_matches_oracle(inlier, use_elc, m, sampler):
    d = synthetic.make_correspondences(8000, inlier_ratio=inlier, seed=33)
    params = engine.make_params(threshold=0.6, confidence=1.0, max_iters=20000, seed=51, sample_size=m,
                                sampler=sampler, use_elc=use_elc, round_size=4096)
    res = engine.ransac_rigid(d["src"], d["tgt"], params, want_mask=True)
    ref = O.ransac(d["src"], d["tgt"], m=m, sampler=sampler, use_elc=use_elc, thr=0.6, conf=1.0,
                   max_iters=20000, round_size=4096, seed=51, return_mask=True)
    assert res["best_id"] == ref["best_id"] and res["best_count"] == ref["best_count"]
    assert res["iters_run"] == ref["iters_run"] == 20000 and res["n_scored"] == ref["n_passed"]
    assert np.array_equal(res["T"], ref["T"])                      # bit-identical fp64 model
    assert np.array_equal(res["mask"].cpu().numpy(), ref["mask"])
    assert res["refit_count"] == ref["refit_count"]
    assert close_T(res["T_refit"], ref["T_refit"])
    assert metrics.registration_success(res["T_refit"], d["T_gt"])


def test_confidence_exit_matches_oracle():
    d = synthetic.make_correspondences(10000, inlier_ratio=0.08, seed=44)
    for conf, R in ((0.9995, 2048), (0.999, 1024), (0.9, 512)):
        params = engine.make_params(confidence=conf, max_iters=200000, seed=7, use_elc=True, round_size=R)
        res = engine.ransac_rigid(d["src"], d["tgt"], params)
        ref = O.ransac(d["src"], d["tgt"], conf=conf, max_iters=200000, round_size=R, seed=7)
        assert res["iters_run"] == ref["iters_run"] < 200000
        assert res["best_id"] == ref["best_id"] and res["best_count"] == ref["best_count"]
        assert np.array_equal(res["T"], ref["T"])


def test_shard_and_finalize_equal_single_call():
    d = synthetic.make_correspondences(5000, inlier_ratio=0.3, seed=55)
    src, tgt = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
    params = engine.make_params(confidence=1.0, max_iters=16384, seed=9, round_size=2048)
    whole = engine.ransac_rigid(src, tgt, params)
    key = torch.zeros(1, dtype=torch.int64, device="cuda")
    for lo, hi in ((8192, 16384), (0, 5000), (5000, 8192)):  # any partition, any order
        engine.ransac_shard(src, tgt, params, lo, hi, key)
    cnt, hid = engine.key_unpack(int(key.item()))
    assert (cnt, hid) == (whole["best_count"], whole["best_id"])
    fin = engine.ransac_finalize(src, tgt, params, int(key.item()))
    assert np.array_equal(fin["T"], whole["T"]) and close_T(fin["T_refit"], whole["T_refit"])


def test_refit_indexed_matches_oracle():
    p = synthetic.make_pair(3000, seed=4, overlap=0.7)
    _, i1, _ = O.find_nn(p["feat0"], p["feat1"])
    i0 = np.arange(3000)
    T, k = engine.refit_indexed(p["xyz0"], p["xyz1"], i0, i1, p["T_gt"], 0.6)
    To, ko = O.refit_indexed(p["xyz0"], p["xyz1"], i0, i1, p["T_gt"], 0.6)
    assert k == ko and close_T(T, To)


def test_degenerate_inputs():
    z = torch.zeros(2, 3)
    res = engine.ransac_rigid(z, z, engine.make_params(max_iters=100))
    assert np.array_equal(res["T"], np.eye(4)) and res["best_id"] == -1
    rng = np.random.default_rng(0)
    a = rng.uniform(-80, 80, (50, 3)).astype(np.float32)
    b = rng.uniform(500, 900, (50, 3)).astype(np.float32)
    res = engine.ransac_rigid(a, b, engine.make_params(threshold=1e-6, use_elc=False, max_iters=50, round_size=64))
    assert np.array_equal(res["T"], np.eye(4))  # zero inliers never replace the identity
    with pytest.raises(RuntimeError):
        engine.ransac_rigid(a, b, engine.make_params(sample_size=5))
    # all points coincide / collinear: finite, orthonormal, equal to the oracle
    c = np.zeros((100, 3), np.float32)
    c[:, 0] = np.arange(100)
    res = engine.ransac_rigid(c, c + 1.0, engine.make_params(use_elc=False, max_iters=256, round_size=256, seed=2))
    ref = O.ransac(c, c + 1.0, use_elc=False, max_iters=256, round_size=256, seed=2)
    assert np.array_equal(res["T"], ref["T"]) and res["best_count"] == ref["best_count"] == 100


def test_large_coordinates_force_fp64_recount():
    """Offsets of 10 km make the fp32 bracket wide: counts must still be exact."""
    d = synthetic.make_correspondences(4000, inlier_ratio=0.5, seed=66)
    src = d["src"] + np.float32(10000.0)
    tgt = d["tgt"] + np.float32(-7000.0)
    rng = np.random.default_rng(2)
    samples = rng.integers(0, 4000, (3000, 3)).astype(np.int32)
    counts, best, _ = engine.ransac_score_samples(src, tgt, samples, 0.6, True, 0.9)
    oc, ob = O.score_samples(src, tgt, samples, 0.6, True, 0.9)
    assert np.array_equal(counts.cpu().numpy(), oc) and best == ob
    params = engine.make_params(max_iters=3000, round_size=1024, seed=5)
    try:
        engine.ransac_set_mode(2)  # the fp32 sweep's bracket is what these offsets widen
        res = engine.ransac_rigid(src, tgt, params)
        assert res["n_rechecked"] > 0
        engine.ransac_set_mode(0)  # the tensor-core sweep works in a frame near the data: same selection
        res_tc = engine.ransac_rigid(src, tgt, params)
        assert res_tc["best_id"] == res["best_id"] and res_tc["best_count"] == res["best_count"]
    finally:
        engine.ransac_set_mode(0)


def test_full_budget_property_1M():
    """cfg 3 at full size: 1M hypotheses on 30k correspondences; checked through properties."""
    d = synthetic.make_correspondences(30000, inlier_ratio=0.3, seed=51 + 3000)
    params = engine.make_params(confidence=1.0, max_iters=1000000, seed=51, use_elc=True)
    res = engine.ransac_rigid(d["src"], d["tgt"], params, want_mask=True)
    assert res["iters_run"] == 1000000 and res["best_count"] > 0.28 * 30000
    # the reported count is the oracle's count of the reported model, and the id regenerates it
    assert res["best_count"] == O.count_inliers(d["src"], d["tgt"], res["T"], 0.6)
    s = O.sample(51, res["best_id"], 0, 3, 30000)
    assert np.array_equal(res["T"], O.kabsch(d["src"][s].astype(float), d["tgt"][s].astype(float)))
    assert int(res["mask"].sum()) == res["best_count"] == res["refit_count"]
    assert metrics.rotation_error_deg(res["T_refit"], d["T_gt"]) < 0.1
    assert metrics.translation_error_cm(res["T_refit"], d["T_gt"]) < 2.0


def test_batch_equals_single_calls():
    """lr_ransac_rigid_batch (two pairs in flight on internal streams) == one lr_ransac_rigid per pair (selection,
    counts and model bit for bit, the least-squares refit within the north star's tolerance);
    pairs of different sizes, a degenerate one (n < sample size) and the confidence exit included."""
    sets = [synthetic.make_correspondences(n, r, seed=400 + k) for k, (n, r) in
            enumerate([(3000, 0.3), (12000, 0.5), (2, 0.5), (7000, 0.1), (9000, 0.6)])]
    pairs = [(torch.from_numpy(d["src"]).cuda(), torch.from_numpy(d["tgt"]).cuda()) for d in sets]
    for kw in (dict(confidence=1.0, max_iters=60000), dict(confidence=0.999, max_iters=200000, round_size=8192)):
        params = engine.make_params(threshold=0.6, seed=7, use_elc=True, **kw)
        single = [engine.ransac_rigid(a, b, params) for a, b in pairs]
        batch = engine.ransac_rigid_batch(pairs, params)
        assert len(batch) == len(single)
        for s1, b1 in zip(single, batch):
            for key in ("best_id", "best_count", "iters_run", "n_scored", "refit_count"):
                assert s1[key] == b1[key], key
            assert np.array_equal(s1["T"], b1["T"])
            assert np.array_equal(s1["T_refit"], b1["T_refit"])  # fixed-order reduction in k_finish: bit-reproducible
    assert engine.ransac_rigid_batch([], params) == []
    # the same pairs handed over in HOST memory (pinned tensors, pageable numpy arrays, and a mix with device tensors):
    # the library stages them with the copy engine on the pair's lane; results identical
    pinned = [(torch.from_numpy(d["src"]).pin_memory(), torch.from_numpy(d["tgt"]).pin_memory()) for d in sets]
    pageable = [(d["src"], d["tgt"]) for d in sets]
    mixed = [(pinned[k][0], pairs[k][1]) if k % 2 else (pairs[k][0], pageable[k][1]) for k in range(len(sets))]
    for host_pairs in (pinned, pageable, mixed):
        hb = engine.ransac_rigid_batch(host_pairs, params)
        for b1, h1 in zip(batch, hb):
            assert all(b1[key] == h1[key] for key in ("best_id", "best_count", "iters_run", "n_scored", "refit_count"))
            assert np.array_equal(b1["T"], h1["T"]) and np.array_equal(b1["T_refit"], h1["T_refit"])


def test_find_rigid_transform_mask_in_pinned_host_memory():
    """The reference-facing call has its inlier mask written by the kernel straight into pinned host memory:
    same mask as the device path, pose = the refit in pygcransac's row-vector convention."""
    from lidarregistration_b200.algorithms import findRigidTransform
    d = synthetic.make_correspondences(8000, 0.4, seed=77)
    pose, mask = findRigidTransform(d["src"], d["tgt"], threshold=0.6, conf=1.0, spatial_coherence_weight=0.0,
                                    max_iters=30000, use_sprt=True, min_inlier_ratio_for_sprt=-1, sampler=0,
                                    neighborhood=0, neighborhood_size=20, seed=5)
    params = engine.make_params(threshold=0.6, confidence=1.0, max_iters=30000, seed=5, use_elc=True, refit=True)
    res = engine.ransac_rigid(d["src"], d["tgt"], params, want_mask=True)
    assert mask.dtype == bool and np.array_equal(mask, res["mask"].cpu().numpy())
    assert int(mask.sum()) == res["best_count"] and close_T(pose.T, res["T_refit"])


def test_sweep_early_out_is_exact():
    """lr_ransac_set_mode: neither the tensor-core sweep (0) nor the warp-uniform early-out of the fp32 sweep (2)
    changes any count of the plain fp32 sweep (1) -- fed samples (every count) and full runs, incl. coordinates
    that force the fp64 recount."""
    d = synthetic.make_correspondences(9000, inlier_ratio=0.4, seed=17)
    rng = np.random.default_rng(5)
    samples = rng.integers(0, 9000, (30000, 3)).astype(np.int32)
    big = {k: (v + np.float32(3000.0) if k in ("src", "tgt") else v) for k, v in d.items()}
    out = {}
    try:
        for mode in (1, 2, 0):
            engine.ransac_set_mode(mode)
            c, b, _ = engine.ransac_score_samples(d["src"], d["tgt"], samples, 0.6, False, 0.9)
            cb, bb, _ = engine.ransac_score_samples(big["src"], big["tgt"], samples[:5000], 0.6, True, 0.9)
            r = engine.ransac_rigid(d["src"], d["tgt"], engine.make_params(max_iters=50000, seed=2))
            out[mode] = (c.cpu().numpy(), b, cb.cpu().numpy(), bb, r["best_id"], r["best_count"], r["n_scored"])
    finally:
        engine.ransac_set_mode(0)
    for x, y, z in zip(out[0], out[1], out[2]):
        assert np.array_equal(x, y) and np.array_equal(x, z)
    oc, ob = O.score_samples(d["src"], d["tgt"], samples[:3000], 0.6, False, 0.9)
    assert np.array_equal(out[0][0][:3000], oc)


def test_pinned_host_inputs_are_read_in_place():
    """engine.ransac_rigid hands pinned fp32 host tensors to the library as they are (k_pack reads them over the
    bus, everything after works on packed device copies): same results as device-resident inputs, both scorings."""
    d = synthetic.make_correspondences(7000, inlier_ratio=0.3, seed=23)
    hs, ht = torch.from_numpy(d["src"]).pin_memory(), torch.from_numpy(d["tgt"]).pin_memory()
    ds, dt = hs.cuda(), ht.cuda()
    for kw in (dict(), dict(scoring=engine.SCORE_MSAC, lo_rounds=3, lo_trials=8, lsq_iters=2)):
        p = engine.make_params(max_iters=40000, seed=4, **kw)
        a = engine.ransac_rigid(hs, ht, p, want_mask=True, mask_on_host=True)
        b = engine.ransac_rigid(ds, dt, p, want_mask=True)
        for k in ("best_id", "best_count", "n_scored", "refit_count", "best_score", "lo_score", "lo_improved"):
            assert a[k] == b[k], k
        # (the least-squares passes of the MSAC run reduce with fp64 atomics: equal to ~1e-15, not bit for bit)
        assert np.abs(a["T"] - b["T"]).max() < 1e-9 and np.array_equal(a["mask"], b["mask"].cpu().numpy())
        if not kw:
            assert np.array_equal(a["T"], b["T"])
        assert close_T(a["T_refit"], b["T_refit"])

