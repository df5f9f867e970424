"""Pipeline trace of the tensor-core inlier sweep (csrc/lr_score_tc.cuh, -DLR_TCS_TRACE): rebuilds the library with
the trace switch, runs one cfg-3 pair, prints per stage the clock64() stamps of CTA 0's MMA warp and epilogue warps.
usage: python tools/tcs_trace.py [TN] [elc 0|1] > profiles/..."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TN = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ELC = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
if os.environ.get("TCS_TRACE_CHILD") != "1":
    flags = "-DLR_TCS_TRACE -DLR_TCS_TN=%d %s" % (TN, os.environ.get("TCS_EXTRA", ""))
    env = dict(os.environ, LIDARREG_NVCC_FLAGS=flags)
    subprocess.run([sys.executable, "-m", "lidarregistration_b200.build", "--force"], cwd=ROOT, env=env, check=True)
    subprocess.run([sys.executable, __file__] + sys.argv[1:], env=dict(os.environ, TCS_TRACE_CHILD="1"), check=False)
    subprocess.run([sys.executable, "-m", "lidarregistration_b200.build", "--force"], cwd=ROOT,
                   env={k: v for k, v in os.environ.items() if k != "LIDARREG_NVCC_FLAGS"}, check=True)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from lidarregistration_b200 import _lib, engine, synthetic  # noqa: E402

d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
a, b = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
p = engine.make_params(threshold=0.6, confidence=1.0, max_iters=1000000 if ELC else 200000, seed=51, use_elc=ELC)
for _ in range(3):
    r = engine.ransac_rigid(a, b, p)
torch.cuda.synchronize()
NT = 128 // TN
LEN = 24
buf = np.zeros((17 + NT, LEN, NT, 4), dtype=np.int64)
rc = _lib.lib().lr_debug_tcs_trace(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.nbytes))
assert rc == 0, rc
mma = np.stack([buf[17 + j][:, j, :] for j in range(NT)], axis=1)  # [LEN, NT, 4]: MMA warp 17 + j issues sub-tile j
t0 = mma[mma > 0].min()
rel = lambda x: int(x - t0) if x > 0 else -1  # noqa: E731
print("# TN %d, ELC %s; cycles relative to the first traced MMA-warp stamp; CTA 0, stages 160..183" % (TN, ELC))
print("# mma (warp 17 + tile): per tile  [before the b_full / t_empty waits, after, issued + committed]   epilogue warp w (quadrant w%4): [before t_full wait, after, "
      "loaded+released, computed]")
for s in range(LEN):
    for j in range(NT):
        print("stage %3d tile %d  mma %s" % (160 + s, j, [rel(x) for x in mma[s][j][:3]]))
        rows = []
        for w in range(16):
            if buf[w][s][j].max() > 0:
                rows.append("w%02d %s" % (w, [rel(x) for x in buf[w][s][j]]))
        for k in range(0, len(rows), 4):
            print("      " + "  ".join(rows[k:k + 4]))
# summary: average period per stage, per-warp phase durations
m = mma[:, 0, 2]
print("# MMA warp: mean cycles per stage %.0f" % ((m[-1] - m[0]) / (LEN - 1)))
w_wait = np.mean([(buf[w][:, :, 1] - buf[w][:, :, 0])[buf[w][:, :, 0] > 0].mean() for w in range(16)])
w_load = np.mean([(buf[w][:, :, 2] - buf[w][:, :, 1])[buf[w][:, :, 0] > 0].mean() for w in range(16)])
w_comp = np.mean([(buf[w][:, :, 3] - buf[w][:, :, 2])[buf[w][:, :, 0] > 0].mean() for w in range(16)])
print("# epilogue warps, mean per tile slice: wait t_full %.0f, load+release %.0f, compute %.0f" % (w_wait, w_load, w_comp))
mw = (mma[:, :, 1] - mma[:, :, 0]).mean()
mi = (mma[:, :, 2] - mma[:, :, 1]).mean()
print("# MMA warps, mean per tile: wait b_full + t_empty %.0f, issue + commits %.0f" % (mw, mi))
full_lat = np.mean([(buf[w][:, :, 1][buf[w][:, :, 0] > 0] - mma[:, :, 2][buf[w][:, :, 0] > 0]).mean() for w in range(16)])
print("# issue -> t_full seen by the epilogue warps (mean over warps and tiles): %.0f" % full_lat)
