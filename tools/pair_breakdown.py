"""Where one cfg-3 pair's time goes on one GPU: library CUDA events around reset + pack, gen + Kabsch, the sweep, the
round end, the finish kernel, against the wall time of the call (host launch + synchronisation included).
usage: python tools/pair_breakdown.py [elc 0|1] [world]   (world > 1: rank 0's slice of a hypothesis-sharded run, no exchange)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import engine, synthetic  # noqa: E402

elc = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if world > 1:
    import ctypes
    from lidarregistration_b200 import _lib
    _lib.check(_lib.lib().lr_debug_slice(0, world), "lr_debug_slice")
d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
a, b = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
p = engine.make_params(threshold=0.6, confidence=1.0, max_iters=1000000, seed=51, use_elc=elc)
for _ in range(5):
    engine.ransac_rigid(a, b, p)
reps = 50
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    engine.ransac_rigid(a, b, p)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / reps * 1e3
kinds = dict(pack=engine.PROF_PACK, gen=engine.PROF_GEN, sweep=engine.PROF_SCORE, round_end=engine.PROF_END, finish=engine.PROF_FIN)
for k in kinds.values():
    engine.prof_read(k)
engine.prof_enable(True)
for _ in range(reps):
    engine.ransac_rigid(a, b, p)
engine.prof_enable(False)
out = {"wall_ms_per_pair": wall, "elc": elc, "world": world}
tot = 0.0
for name, k in kinds.items():
    ms, n = engine.prof_read(k)
    out[name + "_ms"] = ms / reps
    tot += ms / reps
out["sum_of_kernels_ms"] = tot
out["host_and_gaps_ms"] = wall - tot
print(json.dumps(out, indent=1))
