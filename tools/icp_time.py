"""wall time of registration_icp on the 25k-point pair of bench.py's f4 block"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import engine, synthetic  # noqa: E402
from lidarregistration_b200.algorithms import registration_icp  # noqa: E402

p = synthetic.make_pair(25000, seed=51 + 5000, overlap=0.6)
ang = np.deg2rad(1.0)
d = np.eye(4)
d[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
d[:3, 3] = [0.2, -0.15, 0.05]
T0 = d @ p["T_gt"]
a, b = engine.to_dev_f32(p["xyz0"]), engine.to_dev_f32(p["xyz1"])
for _ in range(3):
    r = registration_icp(a, b, 0.6, T0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    r = registration_icp(a, b, 0.6, T0)
torch.cuda.synchronize()
print("icp_ms %.4f iterations %d fitness %.5f rmse %.9f" % ((time.perf_counter() - t0) / 30 * 1e3, r.iterations, r.fitness, r.inlier_rmse))
