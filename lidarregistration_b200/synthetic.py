"""Synthetic LiDAR-shaped registration pairs.

Stands in for the reference's balanced-pair data loader + FCGF network
(Experiments/dataloader/generic_balanced_loader.py:32-98 voxelises scans at
0.3 m; Experiments/datasets/LidarFeatureExtractor.py:166-200 emits unit-norm
32-d features), neither of which can run without the raw datasets and
MinkowskiEngine.  Seeds follow `seed = 51 + 1000*config + pair`
(Experiments/test.py:357 uses seed 51).
"""
import numpy as np

VOXEL = 0.3


def _scan(rng, n):
    """n unique voxel-centre points of a LiDAR-like scan (float64)."""
    pts = np.empty((0, 3))
    keys = np.empty((0,), np.int64)
    while pts.shape[0] < n:
        k = int((n - pts.shape[0]) * 1.6) + 64
        r = 2.0 * 40.0 ** rng.random(k)  # density ~ 1/r on [2, 80] m
        az = rng.random(k) * 2.0 * np.pi
        ground = rng.random(k) < 0.6
        z = np.where(ground, -1.7, rng.uniform(-1.7, 4.0, k))
        p = np.stack([r * np.cos(az), r * np.sin(az), z], axis=1)
        q = np.floor(p / VOXEL).astype(np.int64)
        key = (q[:, 0] + 4096) * (8192 * 8192) + (q[:, 1] + 4096) * 8192 + (q[:, 2] + 4096)
        allk = np.concatenate([keys, key])
        _, first = np.unique(allk, return_index=True)
        first.sort()
        allp = np.concatenate([pts, (q + 0.5) * VOXEL])
        pts, keys = allp[first], allk[first]
    return pts[:n]


def random_motion(rng, yaw_deg=45.0, rp_deg=1.0, txy=30.0, tz=0.3):
    """Balanced-set-like rigid motion; 4x4 float64, column-vector convention."""
    yaw = np.deg2rad(rng.uniform(-yaw_deg, yaw_deg))
    roll, pitch = np.deg2rad(rng.normal(0.0, rp_deg, 2))
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    Ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
    Rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = [rng.uniform(-txy, txy), rng.uniform(-txy, txy), rng.normal(0.0, tz)]
    return T


def _unit(x):
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def make_pair(N, M=None, seed=51, overlap=None, sigma_f=0.08, dim=32, noise=0.05):
    """A full pair: scans + FCGF-shaped features + ground truth.

    -> dict(xyz0[N,3] f32, xyz1[M,3] f32, feat0[N,dim] f32, feat1[M,dim] f32, T_gt[4,4] f64)
    Target = rigid image of an `overlap` fraction of the source + N(0, noise) +
    fresh scan points; overlapping points get f_tgt = normalize(f_src + sigma_f*N(0,I)).
    """
    rng = np.random.default_rng(seed)
    M = N if M is None else M
    xyz0 = _scan(rng, N)
    T = random_motion(rng)
    rho = rng.uniform(0.2, 1.0) if overlap is None else overlap
    k = min(int(round(rho * N)), M)
    sel = rng.permutation(N)[:k]
    moved = xyz0[sel] @ T[:3, :3].T + T[:3, 3] + rng.normal(0.0, noise, (k, 3))
    fresh = _scan(rng, M - k) if M > k else np.empty((0, 3))
    xyz1 = np.concatenate([moved, fresh])
    feat0 = _unit(rng.standard_normal((N, dim)))
    f_ov = _unit(feat0[sel] + sigma_f * rng.standard_normal((k, dim)))
    f_fr = _unit(rng.standard_normal((M - k, dim))) if M > k else np.empty((0, dim))
    feat1 = np.concatenate([f_ov, f_fr])
    perm = rng.permutation(M)  # do not leave the overlap at the front of the target
    return dict(xyz0=xyz0.astype(np.float32), xyz1=xyz1[perm].astype(np.float32),
                feat0=feat0.astype(np.float32), feat1=feat1[perm].astype(np.float32), T_gt=T)


def make_correspondences(n, inlier_ratio=0.3, seed=51, noise=0.1):
    """Direct correspondences (cfg 3 / cfg 4 of BASELINE.json).

    -> dict(src[n,3] f32, tgt[n,3] f32, T_gt, is_inlier[n] bool); inliers are
    T_gt*src + N(0, noise), outliers are random target-scan points.
    """
    rng = np.random.default_rng(seed)
    src = _scan(rng, n)
    T = random_motion(rng)
    tgt = src @ T[:3, :3].T + T[:3, 3] + rng.normal(0.0, noise, (n, 3))
    is_in = rng.random(n) < inlier_ratio
    other = _scan(rng, n) @ T[:3, :3].T + T[:3, 3]
    tgt = np.where(is_in[:, None], tgt, other[rng.permutation(n)])
    return dict(src=src.astype(np.float32), tgt=tgt.astype(np.float32), T_gt=T, is_inlier=is_in)
