"""Drop-in for Experiments/algorithms/FR.py (reference :16-139): the algorithm interface.

FR(A, B, A_feat, B_feat, args, T_gt) keeps the reference's signature, the
8-tuple it returns (FR.py:119), the mode / codebase switches and their
assertion behaviour, and the timing bookkeeping (filter + algorithm + the
extra cost of the 2nd nearest neighbour, FR.py:116-117).  `--mode MMN`
(README.md:55 of the reference, a typo its own code rejects) is accepted as an
alias of `MNN`.  All arithmetic runs in liblidarreg.so.
"""
from copy import deepcopy
from time import time

import numpy as np
import torch

from .. import engine
from .GC_RANSAC import GC_RANSAC, gc_options
from .matching import (Grid_Prioritized_Filter, Grid_Prioritized_Filter_dev, calc_distance_ratio_in_feature_space,  # noqa: F401
                       find_2nn, find_2nn_dev, measure_inlier_ratio, measure_inlier_ratio_dev, nn_to_mutual, nn_to_mutual_dev)

__doc__ = (__doc__ or "") + "\nFast RANSAC algorithms\n"

VOXEL_SIZE = 0.3  # FR.py:18


class PointCloud:
    """Minimal stand-in for o3d.geometry.PointCloud: what the path itself touches
    (`.points`, `.transform(T)`, deepcopy; matching.py:244-247, FR.py:102-104).  When
    open3d is importable the real class is returned instead, so the caller's ICP step
    (Experiments/test.py:185-187) keeps working."""

    def __init__(self, xyz=None):
        self.points = np.zeros((0, 3)) if xyz is None else np.asarray(xyz, dtype=np.float64)

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self.points = self.points @ T[:3, :3].T + T[:3, 3]
        return self

    def __deepcopy__(self, memo):
        return PointCloud(self.points.copy())


def make_point_cloud(xyz):
    try:
        import open3d as o3d  # noqa: F401  (optional: absent in the build image)
        pcd = o3d.geometry.PointCloud()
        pcd.points = o3d.utility.Vector3dVector(xyz)
        return pcd
    except Exception:
        return PointCloud(xyz)


def FR(A, B, A_feat, B_feat, args, T_gt):
    voxel_size = VOXEL_SIZE
    xyz0, xyz1 = A, B
    xyz0_np = xyz0.detach().cpu().numpy().astype(np.float64)
    xyz1_np = xyz1.detach().cpu().numpy().astype(np.float64)
    pcd0 = make_point_cloud(xyz0_np)
    pcd1 = make_point_cloud(xyz1_np)

    fcgf_feats0 = engine.to_dev_f32(A_feat)  # FR.py:32-34 [H->D]
    fcgf_feats1 = engine.to_dev_f32(B_feat)
    xyz0_d, xyz1_d = engine.to_dev_f32(xyz0), engine.to_dev_f32(xyz1)
    mode = "MNN" if args.mode == "MMN" else args.mode

    # The index tensors stay in HBM from the sweep to the RANSAC call (the reference moves them to the CPU after
    # every step, matching.py:62-65); only counts come back to the host (--mode GPF included: csrc/lr_gpf.cu).
    with torch.no_grad():
        # 1. Coarse correspondences
        corres_idx1, idx1_2nd, additional_time_for_finding_2nd_closest = find_2nn_dev(fcgf_feats0, fcgf_feats1)
        corres_idx0 = torch.arange(corres_idx1.shape[0], device=corres_idx1.device)
        num_pairs_init = len(corres_idx0)
        inlier_ratio_init = measure_inlier_ratio_dev(corres_idx0, corres_idx1, xyz0_d, xyz1_d, T_gt, voxel_size)

        torch.cuda.synchronize()
        start_time = time()
        # 2. Filter correspondences
        norm_feat_dist = None
        if mode == "MNN":
            corres_idx0_orig, corres_idx1_orig = corres_idx0, corres_idx1
            corres_idx0, corres_idx1, idx1_2nd = nn_to_mutual_dev(fcgf_feats0, fcgf_feats1, corres_idx1, idx1_2nd)
        elif mode == "GPF":
            corres_idx0, corres_idx1, idx1_2nd, corres_idx0_orig, corres_idx1_orig, _, norm_feat_dist = \
                Grid_Prioritized_Filter_dev(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd, xyz0_d, args)
        elif mode == "no_filter":
            corres_idx0_orig, corres_idx1_orig = corres_idx0, corres_idx1
        else:
            assert False, "unknown mode"
        torch.cuda.synchronize()
        filter_time = time() - start_time

        num_pairs_filtered = len(corres_idx0)
        inlier_ratio_filtered = measure_inlier_ratio_dev(engine.to_dev_i64(corres_idx0), engine.to_dev_i64(corres_idx1),
                                                         xyz0_d, xyz1_d, T_gt, voxel_size)

    start_time = time()
    ransac_iters = 500 * 10 ** 3  # FR.py:65
    if args.iters is not None:
        ransac_iters = args.iters

    # 3. Perform RANSAC
    if args.codebase == "GC":
        # Same steps as the reference (FR.py:70-87 -> GC_RANSAC.py:8-55): gather the filtered
        # correspondences, PROSAC quality = -ratio, sort best first, native call, None -> identity.
        # Everything stays in HBM: the gathers, the ratio and the (stable) sort run on the device and
        # lr_ransac_rigid is called directly instead of round-tripping through numpy.
        src = engine.gather_xyz(xyz0_d, corres_idx0)
        tgt = engine.gather_xyz(xyz1_d, corres_idx1)
        if args.prosac and num_pairs_filtered > 0:
            if mode == 'GPF':
                feat_dist = engine.to_dev_f32(norm_feat_dist)
            else:
                feat_dist = engine.match_ratio(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd)
            order = torch.argsort(feat_dist, stable=True)  # == argsort(-match_quality), match_quality = -ratio
            src, tgt = src[order].contiguous(), tgt[order].contiguous()
        # --fast_rejection: ELC = edge-length test, NONE = the glue's no-preemption branch, SPRT = every hypothesis
        # scored in full (GC_RANSAC.py docstring: the sequential test is a speed-up device of the reference engine)
        gc = gc_options(getattr(args, "GC_scoring", "count"), args.GC_LO, args.spatial_coherence_weight,
                        preemption=args.fast_rejection != "NONE")
        params = engine.make_params(threshold=2 * voxel_size, confidence=args.GC_conf, max_iters=ransac_iters,
                                    seed=getattr(args, "seed", 51), sample_size=3,
                                    sampler=engine.SAMPLER_PROSAC if args.prosac else engine.SAMPLER_UNIFORM,
                                    use_elc=args.fast_rejection == "ELC", elc_ratio=0.9, refit=True, **gc)
        res = engine.ransac_rigid(src, tgt, params)
        # count scoring ends with the least-squares refit over the winner's inliers; the MSAC run's own
        # local optimisation + iterated least squares already produced the final model (GC_RANSAC.py docstring)
        T = res["T" if gc["scoring"] == engine.SCORE_MSAC else "T_refit"] if res["best_count"] > 0 else np.eye(4)

    elif args.codebase == "open3D":
        T = RANSAC_registration(pcd0, pcd1, corres_idx0, corres_idx1, 2 * voxel_size, num_iterations=ransac_iters,
                                args=args)
        # estimate motion using all inlier pairs of the ORIGINAL (unfiltered) NN set (FR.py:99-111)
        T, _ = engine.refit_indexed(xyz0_d, xyz1_d, corres_idx0_orig, corres_idx1_orig, T, 2 * voxel_size)
    else:
        assert False, "unknown codebase"

    torch.cuda.synchronize()
    algo_time = time() - start_time
    elapsed_time = filter_time + algo_time + additional_time_for_finding_2nd_closest
    return T, elapsed_time, pcd0, pcd1, num_pairs_init, inlier_ratio_init, num_pairs_filtered, inlier_ratio_filtered


def RANSAC_registration(pcd0, pcd1, idx0, idx1, distance_threshold, num_iterations, args):
    """FR.py:122-139: Open3D registration_ransac_based_on_correspondence with ransac_n=4,
    CorrespondenceCheckerBasedOnEdgeLength (0.9), confidence 0.9995, no refit inside."""
    xyz0 = torch.from_numpy(np.asarray(pcd0.points)).float()
    xyz1 = torch.from_numpy(np.asarray(pcd1.points)).float()
    src = engine.gather_xyz(xyz0, idx0)
    tgt = engine.gather_xyz(xyz1, idx1)
    if src.shape[0] < 4:  # Open3D: |corres| < ransac_n -> identity
        return np.eye(4)
    params = engine.make_params(threshold=distance_threshold, confidence=0.9995, max_iters=num_iterations,
                                seed=getattr(args, "seed", 51), sample_size=4, sampler=engine.SAMPLER_REPLACE,
                                use_elc=True, elc_ratio=0.9, refit=False)
    return engine.ransac_rigid(src, tgt, params)["T"]
