"""Drop-in for Experiments/algorithms/GC_RANSAC.py (reference :8-55).

Keeps the signature `GC_RANSAC(A, B, distance_threshold, num_iterations, args,
match_quality) -> (T[4,4], seconds)` and the flag mapping of the reference
(GC_RANSAC.py:12-43), but the native call is liblidarreg.so's
lr_ransac_rigid instead of pygcransac.findRigidTransform
(gcransac_python.cpp:404-624).  The reference file does not byte-compile under
Python 3 (tab-indented lines 29-30); this one does.

What the flags select here (DESIGN.md "GC codebase mapping"):
  --fast_rejection ELC  -> edge-length pre-rejection (preemption_edge_length.h:71-128)
  --fast_rejection NONE -> no pre-rejection
  --fast_rejection SPRT -> not implemented (raises)
  --prosac True         -> correspondences pre-sorted best-first (GC_RANSAC.py:39-43) and the
                           PROSAC progressive sampler (sampler id 1, gcransac_python.cpp:464-465)
  --GC_conf c           -> confidence of the stopping rule
  --GC_LO               -> graph-cut LO is SURVEY row f3; the final least-squares
                           refit over the inliers is always returned, as pygcransac does
"""
from time import time

import numpy as np
import torch

from .. import engine


def findRigidTransform(x1y1z1, x2y2z2, threshold, conf, spatial_coherence_weight, max_iters, use_sprt,
                       min_inlier_ratio_for_sprt, sampler, neighborhood, neighborhood_size, seed=51,
                       round_size=engine.DEFAULT_ROUND):
    """Same kwargs as pygcransac.findRigidTransform (GC_RANSAC.py:12-22).

    -> (pose[4,4] float64 in pygcransac's ROW-vector convention | None, mask[n] bool)
    """
    if use_sprt and not (min_inlier_ratio_for_sprt < 0):
        raise NotImplementedError("SPRT pre-verification is not part of the B200 hot path (use ELC or NONE)")
    params = engine.make_params(threshold=threshold, confidence=conf, max_iters=max_iters, seed=seed, sample_size=3,
                                sampler=engine.SAMPLER_PROSAC if int(sampler) == 1 else engine.SAMPLER_UNIFORM,
                                use_elc=bool(use_sprt), elc_ratio=0.9,
                                round_size=round_size, refit=True)
    res = engine.ransac_rigid(x1y1z1, x2y2z2, params, want_mask=True, mask_on_host=True)
    mask = res["mask"]
    if res["best_count"] <= 0:  # 0 inliers: Python gets None (gcransac_python.cpp:594-611)
        return None, mask
    return res["T_refit"].T.copy(), mask


def GC_RANSAC(A, B, distance_threshold, num_iterations, args, match_quality):
    x1y1z1_ = np.ascontiguousarray(A)
    x2y2z2_ = np.ascontiguousarray(B)
    params = {
        'threshold': distance_threshold,
        'conf': args.GC_conf,
        'spatial_coherence_weight': args.spatial_coherence_weight,
        'max_iters': num_iterations,
        'use_sprt': args.fast_rejection != "NONE",  # "perform fast rejection"
        'min_inlier_ratio_for_sprt': -1 if args.fast_rejection == "ELC" else 0.1,  # < 0 selects ELC
        'sampler': args.prosac,
        'neighborhood': 0 if args.GC_LO else 1,  # non-zero: no local optimisation
        'neighborhood_size': 20,
    }
    if args.prosac:
        order = np.argsort(-match_quality, kind="stable")  # best quality first
        x1y1z1_ = x1y1z1_[order, :]
        x2y2z2_ = x2y2z2_[order, :]

    torch.cuda.synchronize()
    start_time = time()
    pose_T, mask = findRigidTransform(x1y1z1_, x2y2z2_, seed=getattr(args, "seed", 51), **params)
    if pose_T is None:
        pose_T = np.eye(4, dtype=np.float32)
    elapsed_time = time() - start_time
    return pose_T.T, elapsed_time
