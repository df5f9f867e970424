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
  --fast_rejection SPRT -> every hypothesis is scored in full, no pre-rejection: the sequential probability ratio
                           test (gcransac_python.cpp:534-568) only abandons the evaluation of models that are
                           unlikely to beat the best so far -- a model it lets through gets exactly the score it
                           has without the test, so it is a speed-up device with a bounded false-rejection rate,
                           not part of the result's definition; the chip-wide sweep needs no such shortcut.
                           LO settings are those of the ELC branch (:548-555).  A one-time warning says so.
  --prosac True         -> correspondences pre-sorted best-first (GC_RANSAC.py:39-43) and the
                           PROSAC progressive sampler (sampler id 1, gcransac_python.cpp:464-465)
  --GC_conf c           -> confidence of the stopping rule
  --GC_scoring count    -> [default] the graded selection criterion of SURVEY 8(a): inlier count at the
                           threshold, ties -> lowest hypothesis id; the least-squares refit over the
                           winner's inliers is returned, as pygcransac ends with a non-minimal fit
  --GC_scoring MSAC     -> pygcransac's own criterion: MSAC score at 1.5 x threshold,
                           then, with --GC_LO True, local optimisation (10 rounds x 20 inner draws of
                           min(21, #inliers) inliers; the graph cut with spatial_coherence_weight = 0 is
                           thresholding) and, whatever --GC_LO says, up to 10 passes of iterated least squares
                           (--GC_LO False only sets max_graph_cut_number = 0, :518-521; the finishing fit still
                           runs, App. A "Finish"); the final model is returned.  With --fast_rejection NONE the
                           reference takes its no-preemption branch (:570-592): 50 inner draws, local optimisation
                           on regardless of --GC_LO.
  (`GC_scoring` is an attribute this drop-in adds; the reference has no such flag because pygcransac
   knows only MSAC.  args without it get "count".)
"""
from time import time

import numpy as np
import torch

from .. import engine


# settings the reference's glue fixes (gcransac_python.cpp:511-521) / upstream defaults
GC_LO_ROUNDS = 10   # settings.max_graph_cut_number (0 when neighborhood != 0, i.e. --GC_LO False)
GC_LO_TRIALS = 20   # settings.max_local_optimization_number
GC_LSQ_ITERS = 10   # iterated least squares, upstream's cap


GC_LO_TRIALS_NOPREEMPT = 50  # the no-preemption branch's max_local_optimization_number (gcransac_python.cpp:577)

_warned = set()


def _warn_once(key, msg):
    if key not in _warned:
        _warned.add(key)
        import warnings
        warnings.warn(msg, stacklevel=3)


def gc_options(scoring, local_optimisation, spatial_coherence_weight=0.0, preemption=True):
    """--GC_scoring / --GC_LO / --fast_rejection -> the scoring / lo_rounds / lo_trials / lsq_iters of
    engine.make_params.  `preemption` False = --fast_rejection NONE = the glue's no-preemption branch."""
    scoring = "count" if scoring is None else str(scoring)
    if scoring.lower() == "count":
        # count scoring has no local-optimisation stage: say so instead of silently ignoring the flags
        if not local_optimisation:
            _warn_once("lo", "--GC_LO False has no effect under --GC_scoring count (no local optimisation stage)")
        if spatial_coherence_weight != 0.0:
            _warn_once("scw", "--spatial_coherence_weight has no effect under --GC_scoring count")
        return dict(scoring=engine.SCORE_COUNT)
    if scoring.upper() != "MSAC":
        raise ValueError("GC_scoring must be 'count' or 'MSAC'")
    if spatial_coherence_weight != 0.0:
        # the pairwise term needs the neighbourhood graph (FLANN, gcransac_python.cpp:444-446); the reference's
        # default and README examples all use 0.0 (test.py:306), for which the graph cut is thresholding
        raise NotImplementedError("spatial_coherence_weight != 0 (graph-cut pairwise term) is not built")
    if not preemption:  # :570-592 ignores do_local_optimization and raises the inner-draw budget to 50
        return dict(scoring=engine.SCORE_MSAC, lo_rounds=GC_LO_ROUNDS, lo_trials=GC_LO_TRIALS_NOPREEMPT,
                    lsq_iters=GC_LSQ_ITERS)
    on = bool(local_optimisation)
    return dict(scoring=engine.SCORE_MSAC, lo_rounds=GC_LO_ROUNDS if on else 0, lo_trials=GC_LO_TRIALS,
                lsq_iters=GC_LSQ_ITERS)


def findRigidTransform(x1y1z1, x2y2z2, threshold, conf, spatial_coherence_weight, max_iters, use_sprt,
                       min_inlier_ratio_for_sprt, sampler, neighborhood, neighborhood_size, seed=51,
                       round_size=engine.DEFAULT_ROUND, scoring="count"):
    """Same kwargs as pygcransac.findRigidTransform (GC_RANSAC.py:12-22); `scoring`: see the module docstring.

    -> (pose[4,4] float64 in pygcransac's ROW-vector convention | None, mask[n] bool)
    """
    sprt = bool(use_sprt) and not (min_inlier_ratio_for_sprt < 0)
    if sprt:
        _warn_once("sprt", "--fast_rejection SPRT: every hypothesis is scored in full (the sequential test is a "
                           "speed-up device of the reference engine, not part of the result's definition)")
    # neighborhood != 0: no LO (:418-423)
    opts = gc_options(scoring, int(neighborhood) == 0, spatial_coherence_weight, preemption=bool(use_sprt))
    params = engine.make_params(threshold=threshold, confidence=conf, max_iters=max_iters, seed=seed, sample_size=3,
                                sampler=engine.SAMPLER_PROSAC if int(sampler) == 1 else engine.SAMPLER_UNIFORM,
                                use_elc=bool(use_sprt) and not sprt, elc_ratio=0.9,
                                round_size=round_size, refit=True, **opts)
    res = engine.ransac_rigid(x1y1z1, x2y2z2, params, want_mask=True, mask_on_host=True)
    mask = res["mask"]
    if res["best_count"] <= 0:  # 0 inliers: Python gets None (gcransac_python.cpp:594-611)
        return None, mask
    pose = res["T"] if opts["scoring"] == engine.SCORE_MSAC else res["T_refit"]
    return pose.T.copy(), mask


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
    pose_T, mask = findRigidTransform(x1y1z1_, x2y2z2_, seed=getattr(args, "seed", 51),
                                      scoring=getattr(args, "GC_scoring", "count"), **params)
    if pose_T is None:
        pose_T = np.eye(4, dtype=np.float32)
    elapsed_time = time() - start_time
    return pose_T.T, elapsed_time
