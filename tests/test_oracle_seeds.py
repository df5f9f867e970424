"""SURVEY 8(f4), the consumers right after the path, pinned to the reference run in the build container
(tests/golden/seeds_ref.npz <- PointDSC.cal_seed_trans, Experiments/models/PointDSC.py:234-336):
the per-seed weighted Kabsch (models/common.py:7-45) and the seed scoring (:319-336); and the radius nearest
neighbour behind the ICP refinement (Experiments/test.py:183-188) against a plain numpy statement of it."""
import numpy as np

from oracle import lr_oracle as O

# The reference computes in fp32: |T p - q| is rounded (coordinates ~100 m -> ~1e-5 m) before it is compared with the
# threshold, the oracle decides on the fp64 residual of the same fp32 inputs.  Residuals this close to the threshold
# are "ambiguous": the two sides may label them differently; everything else must agree exactly.
AMBIG = 2e-4


def wcost(T, A, B, w):
    return float(np.sum(w * np.sum((A @ T[:3, :3].T + T[:3, 3] - B) ** 2, axis=1)))


def test_weighted_kabsch_against_reference(seeds_golden):
    for g in seeds_golden:
        A, B, w, ref = g["A"], g["B"], g["w"], g["trans"]
        for s in range(len(A)):
            a, b, ww = A[s].astype(np.float64), B[s].astype(np.float64), w[s].astype(np.float64)
            mine = O.kabsch_weighted(A[s], B[s], w[s])
            assert abs(np.linalg.det(mine[:3, :3]) - 1) < 1e-12
            # the rotation is the weighted least-squares optimum: never worse than the reference's fp32 solution
            # (compared on centred coordinates, where the reference's 1e-6 in the centroid denominator plays no role)
            ca, cb = (ww[:, None] * a).sum(0) / ww.sum(), (ww[:, None] * b).sum(0) / ww.sum()
            cm = float(np.sum(ww * np.sum(((a - ca) @ mine[:3, :3].T - (b - cb)) ** 2, axis=1)))
            cr = float(np.sum(ww * np.sum(((a - ca) @ ref[s][:3, :3].astype(np.float64).T - (b - cb)) ** 2, axis=1)))
            assert cm <= cr * (1 + 1e-4) + 1e-7
            if np.sort(np.linalg.svd((ww[:, None] * (a - ca)).T @ (b - cb), compute_uv=False))[1] > 1e-3:
                assert np.allclose(mine[:3, :3], ref[s][:3, :3], atol=5e-3)
                # t = cB - R cA with the reference's own centroids (denominator sum w + 1e-6)
                assert np.allclose(mine[:3, 3], ref[s][:3, 3], atol=5e-3 * (1 + np.abs(ca).max()))


def test_seed_scoring_against_reference(seeds_golden):
    for g in seeds_golden:
        src, tgt, thr = g["src"], g["tgt"], g["threshold"]
        trans = g["trans"].astype(np.float64)
        counts, best, labels = O.seeds_score(src, tgt, trans, thr, return_labels=True)
        n = len(src)
        ref_counts = np.rint(g["fitness"].astype(np.float64) * n).astype(np.int64)
        p, q = src.astype(np.float64), tgt.astype(np.float64)
        for s in range(len(trans)):
            r = np.linalg.norm(p @ trans[s, :3, :3].T + trans[s, :3, 3] - q, axis=1)
            amb = int(np.sum(np.abs(r - thr) < AMBIG))
            assert abs(int(counts[s]) - int(ref_counts[s])) <= amb, (s, counts[s], ref_counts[s], amb)
            assert int(counts[s]) == int(np.sum(r < thr)) or amb > 0
        # the selected seed: the reference's arg-max unless fp32-ambiguous residuals could change the order
        ref_best = int(np.argmax(g["fitness"]))
        assert counts[best] == counts.max() and best == int(np.argmax(counts))
        assert best == ref_best or abs(int(counts[best]) - int(counts[ref_best])) <= 4
        if best == ref_best:
            r = np.linalg.norm(p @ trans[best, :3, :3].T + trans[best, :3, 3] - q, axis=1)
            clear = np.abs(r - thr) >= AMBIG
            assert np.array_equal(labels[clear], g["final_labels"].astype(bool)[clear])
            assert np.allclose(trans[best], g["final_trans"].astype(np.float64))


def test_nn3d_radius_against_numpy():
    rng = np.random.default_rng(5)
    tgt = rng.uniform(-20, 20, (3000, 3)).astype(np.float32)
    src = (tgt[rng.permutation(3000)[:2000]] + rng.normal(0, 0.3, (2000, 3))).astype(np.float32)
    src[:50] += 100.0  # nothing within the radius
    tgt[100] = tgt[7]  # duplicate target points: the lowest index wins
    ang = 0.05
    T = np.eye(4)
    T[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    T[:3, 3] = [0.1, -0.2, 0.05]
    idx, d2 = O.nn3d_radius(src, tgt, T, 0.6)
    moved = src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    D = ((moved[:, None, :] - tgt[None, :, :].astype(np.float64)) ** 2).sum(-1)
    j = D.argmin(1)
    ok = D[np.arange(len(src)), j] < 0.36
    assert np.array_equal(idx >= 0, ok)
    assert np.array_equal(idx[ok], j[ok])
    assert np.allclose(d2[ok], D[np.arange(len(src)), j][ok], rtol=1e-12)
    assert (idx[:50] == -1).all()


def test_nn3d_radius_against_scipy_kdtree():
    """an independent KD-tree (scipy's cKDTree, the role nanoflann plays inside Open3D's registration_icp) returns the
    same neighbour inside the radius for every query of a LiDAR-shaped pair"""
    from scipy.spatial import cKDTree
    from lidarregistration_b200 import synthetic
    p = synthetic.make_pair(6000, seed=123, overlap=0.7)
    T = p["T_gt"].copy()
    T[:3, 3] += [0.1, -0.05, 0.02]
    idx, d2 = O.nn3d_radius(p["xyz0"], p["xyz1"], T, 0.6)
    moved = p["xyz0"].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    dist, j = cKDTree(p["xyz1"].astype(np.float64)).query(moved, k=1, distance_upper_bound=0.6)
    found = np.isfinite(dist)
    # (scipy's bound is inclusive, the oracle's strict; no query sits exactly on the radius here)
    assert np.array_equal(found, idx >= 0)
    same = idx[found] == j[found]
    # a different index is only acceptable for an exact tie in distance
    assert np.allclose(np.sqrt(d2[found][~same]), dist[found][~same], rtol=0, atol=1e-12)
    assert same.mean() > 0.999
    assert np.allclose(np.sqrt(d2[found]), dist[found], atol=1e-9)


def test_icp_against_independent_kdtree_svd_icp():
    """the whole refinement against a second, independent statement of Open3D's loop: scipy cKDTree for the
    correspondences, numpy SVD (R = V diag(1, 1, det) U^T) for the point-to-point update, the same relative
    fitness / rmse stopping rule"""
    from scipy.spatial import cKDTree
    from lidarregistration_b200 import synthetic
    p = synthetic.make_pair(5000, seed=321, overlap=0.8)
    src, tgt = p["xyz0"].astype(np.float64), p["xyz1"].astype(np.float64)
    T0 = p["T_gt"].copy()
    T0[:3, 3] += [0.2, -0.15, 0.05]
    tree = cKDTree(tgt)

    def evaluate(T):
        moved = src @ T[:3, :3].T + T[:3, 3]
        dist, j = tree.query(moved, k=1, distance_upper_bound=0.6)
        keep = np.isfinite(dist)
        a, b = src[keep], tgt[j[keep]]
        ca, cb = a.mean(0), b.mean(0)
        U, _, Vt = np.linalg.svd((a - ca).T @ (b - cb))
        D = np.diag([1.0, 1.0, np.linalg.det(Vt.T @ U.T)])
        R = Vt.T @ D @ U.T
        Tn = np.eye(4)
        Tn[:3, :3], Tn[:3, 3] = R, cb - R @ ca
        return Tn, keep.sum() / len(src), float(np.sqrt((dist[keep] ** 2).mean()))

    T = T0
    Tn, fit, rmse = evaluate(T)
    it = 0
    for it in range(1, 31):
        T = Tn
        Tn, f2, r2 = evaluate(T)
        done = abs(fit - f2) < 1e-6 and abs(rmse - r2) < 1e-6
        fit, rmse = f2, r2
        if done:
            break
    To, fo, ro, ito = O.icp(p["xyz0"], p["xyz1"], 0.6, T0)
    assert ito == it and abs(fo - fit) < 1e-12 and abs(ro - rmse) < 1e-9
    assert np.abs(To - T).max() < 1e-8


def test_icp_oracle_converges():
    from lidarregistration_b200 import synthetic
    p = synthetic.make_pair(4000, 4000, seed=77, overlap=0.7)
    Tgt = p["T_gt"]
    d = np.eye(4)
    d[:3, 3] = [0.15, -0.1, 0.05]
    T0 = d @ Tgt
    T, fit, rmse, it = O.icp(p["xyz0"], p["xyz1"], 0.6, T0)
    assert fit > 0.3 and it >= 1 and 0 < rmse < 0.6
    assert np.linalg.norm(T[:3, 3] - Tgt[:3, 3]) < 0.5 * np.linalg.norm(T0[:3, 3] - Tgt[:3, 3])
    # max_iteration = 0: the initial transform is only evaluated
    T1, f1, r1, it1 = O.icp(p["xyz0"], p["xyz1"], 0.6, T0, max_iteration=0)
    assert it1 == 0 and np.array_equal(T1, T0) and f1 > 0
