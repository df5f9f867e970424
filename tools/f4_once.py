"""one ICP refinement + one seed scoring (ncu target for the f4 kernels)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import engine, synthetic  # noqa: E402
from lidarregistration_b200.algorithms import registration_icp  # noqa: E402

p = synthetic.make_pair(25000, seed=51 + 5000, overlap=0.6)
T0 = p["T_gt"].copy()
T0[:3, 3] += [0.2, -0.15, 0.05]
r = registration_icp(p["xyz0"], p["xyz1"], 0.6, T0)
print("icp", r.iterations, r.fitness)
d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
rng = np.random.default_rng(0)
A, B, w = rng.normal(size=(3000, 40, 3)).astype(np.float32), rng.normal(size=(3000, 40, 3)).astype(np.float32), rng.random((3000, 40)).astype(np.float32)
engine.kabsch_weighted_batch(A, B, w)
models = np.tile(d["T_gt"], (3000, 1, 1))
res = engine.seeds_score(d["src"], d["tgt"], models, 0.6)
print("seeds", res["best"], res["best_count"])
