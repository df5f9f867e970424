"""CPU estimate of the early-out rate of k_score under different orderings of the surviving
hypotheses: fraction of (warp of 64 hypotheses, correspondence) pairs in which some hypothesis has |d0| < c.
Uses the oracle's sampler / ELC / Kabsch, numpy for the first residual component.  usage: python tests/sim_early_out.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidarregistration_b200 import synthetic  # noqa: E402
from oracle import lr_oracle as O  # noqa: E402

n, H, thr = 30000, 200000, 0.6
d = synthetic.make_correspondences(n, 0.3, seed=51 + 3000)
src, tgt = d["src"].astype(np.float64), d["tgt"].astype(np.float64)
rng = np.random.default_rng(0)
models = []
for h in range(H):
    s = O.sample(51, h, O.UNIFORM, 3, n)
    if not O.elc(src[s], tgt[s], 0.9):
        continue
    models.append(O.kabsch(src[s], tgt[s])[:3, :])
M = np.stack(models)  # [S, 3, 4]
S = len(M)
print("survivors", S, "of", H)
P = np.concatenate([src, np.ones((n, 1))], 1)  # [n, 4]


def live_fraction(order, comp=0, warps=60):
    tot = live = 0
    for w in range(min(warps, S // 64)):
        idx = order[w * 64:(w + 1) * 64]
        d0 = M[idx, comp, :] @ P.T - tgt[:, comp][None, :]  # [64, n]
        live += int((np.abs(d0) < thr).any(0).sum())
        tot += n
    return live / tot


ident = np.arange(S)
print("id order, x component      :", round(live_fraction(ident, 0), 4))
print("id order, y component      :", round(live_fraction(ident, 1), 4))
print("id order, z component      :", round(live_fraction(ident, 2), 4))
# similarity key: where the model sends the scene centroid + its yaw
c = np.append(src.mean(0), 1.0)
img = M @ c  # [S, 3]
yaw = np.arctan2(M[:, 1, 0], M[:, 0, 0])
key = np.lexsort((np.round(yaw / 0.01), np.round(img[:, 1] / 0.5), np.round(img[:, 0] / 0.5)))
print("sorted by (x, y, yaw) cells:", round(live_fraction(key, 0), 4))
good = np.abs(img - (d["T_gt"][:3, :] @ c)).max(1) < 1.0
print("fraction of survivors within 1 m of the true motion at the centroid:", round(float(good.mean()), 3))
gi = np.flatnonzero(good)
print("only near-true models      :", round(live_fraction(gi, 0), 4))
bi = np.flatnonzero(~good)
print("only far models            :", round(live_fraction(bi, 0), 4))
print("inlier correspondences     :", round(float(d["is_inlier"].mean()), 3))
