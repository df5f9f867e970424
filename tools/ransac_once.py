"""One cfg-3 RANSAC pair (30k correspondences, 1M hypotheses, ELC on) repeated a few times: the workload ncu
profiles the RANSAC kernels on.  usage: python tools/ransac_once.py [score_mode] [reps] [msac]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import engine, synthetic  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
msac = len(sys.argv) > 3 and sys.argv[3] == "msac"
d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
a, b = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
engine.ransac_set_mode(mode)
kw = dict(scoring=engine.SCORE_MSAC, lo_rounds=10, lo_trials=20, lsq_iters=10) if msac else {}
p = engine.make_params(threshold=0.6, confidence=1.0, max_iters=1000000, seed=51, use_elc=True, **kw)
for _ in range(reps):
    r = engine.ransac_rigid(a, b, p)
torch.cuda.synchronize()
print({k: v for k, v in r.items() if k not in ("T", "T_refit", "mask")})
