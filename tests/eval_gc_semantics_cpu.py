#!/usr/bin/env python
"""CPU-only accuracy comparison of the two selection semantics of the reference on synthetic
LiDAR-shaped pairs, through the ORACLE pipeline (test infrastructure; the CUDA path is bit-exact with it for the
selection and within 1e-9 for the final model, tests/test_gpu_z_gc.py):
  count     inlier count at thr, lowest id, least-squares refit over the winner's inliers   (graded criterion)
  msac      MSAC at 1.5 thr, no polishing                                                    (--GC_LO False)
  msac+lo   MSAC, 10 x 20 local-optimisation draws, 10 least-squares passes                  (--GC_LO True)
RRE / RTE / recall use the reference's definitions (Experiments/libs/loss.py:44-51, 5 deg / 60 cm).

    python tests/eval_gc_semantics_cpu.py --pairs 40 --points 6000 --iters 20000 --out profiles/r1_gc_semantics_accuracy.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidarregistration_b200 import metrics, synthetic  # noqa: E402
from oracle import lr_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=40)
    ap.add_argument("--points", type=int, default=6000)
    ap.add_argument("--iters", type=int, default=20000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    T = dict(count=[], count_raw=[], msac=[], msac_lo=[])
    Tgt, npairs = [], []
    for p in range(a.pairs):
        rng = np.random.default_rng(51 + 7000 + p)
        n = int(a.points * rng.uniform(0.8, 1.2))
        d = synthetic.make_pair(n, seed=51 + 7000 + p, sigma_f=float(rng.uniform(0.08, 0.16)),
                                overlap=float(rng.uniform(0.15, 0.6)))
        _, i1, _ = O.find_nn(d["feat0"], d["feat1"])
        m0, m1 = O.nn_to_mutual(d["feat0"], d["feat1"], i1)
        A, B = d["xyz0"][m0], d["xyz1"][m1]
        kw = dict(m=3, sampler=O.UNIFORM, use_elc=True, thr=0.6, conf=1.0, max_iters=a.iters, seed=51)
        r = O.ransac(A, B, **kw)
        T["count_raw"].append(r["T"] if r["best_count"] > 0 else np.eye(4))
        T["count"].append(r["T_refit"] if r["best_count"] > 0 else np.eye(4))
        T["msac"].append(O.ransac_gc(A, B, lo_rounds=0, lsq_iters=0, **kw)["T"])
        T["msac_lo"].append(O.ransac_gc(A, B, lo_rounds=10, lo_trials=20, lsq_iters=10, **kw)["T"])
        Tgt.append(d["T_gt"])
        npairs.append(len(m0))
    out = dict(pairs=a.pairs, points=a.points, iters=a.iters, mutual_pairs_mean=float(np.mean(npairs)),
               note="oracle pipeline on CPU; 'count_raw' is the selected 3-point model before the refit",
               **{k: metrics.summarize(v, Tgt) for k, v in T.items()})
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
