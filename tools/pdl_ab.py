"""A/B of the programmatic dependent launch along a run's kernel chain (lr_debug_pdl): wall time per cfg-3 pair of the
single-pair call, whole and for rank 0's slice of an 8-rank hypothesis-sharded run (lr_debug_slice), and of the batched entry."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidarregistration_b200 import _lib, engine, synthetic  # noqa: E402

d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
a, b = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
p = engine.make_params(threshold=0.6, confidence=1.0, max_iters=1000000, seed=51, use_elc=True)
L = _lib.lib()
out = {}
ref = None
for world in (1, 8):
    _lib.check(L.lr_debug_slice(0, world), "slice")
    for pdl in (0, 1, 0, 1):
        _lib.check(L.lr_debug_pdl(pdl), "pdl")
        for _ in range(10):
            r = engine.ransac_rigid(a, b, p)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            r = engine.ransac_rigid(a, b, p)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 200 * 1e3
        sig = (r["best_id"], r["best_count"], r["n_scored"], float(r["T_refit"][0, 3]))
        if world == 1:
            ref = ref or sig
            assert sig == ref, (sig, ref)
        out.setdefault("world%d_pdl%d" % (world, pdl), []).append(round(ms, 4))
_lib.check(L.lr_debug_slice(0, 1), "slice")
pairs = [(a, b)] * 16
for pdl in (0, 1, 0, 1):
    _lib.check(L.lr_debug_pdl(pdl), "pdl")
    for _ in range(3):
        engine.ransac_rigid_batch(pairs, p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        rb = engine.ransac_rigid_batch(pairs, p)
    torch.cuda.synchronize()
    out.setdefault("batch16_pdl%d" % pdl, []).append(round((time.perf_counter() - t0) / 160 * 1e3, 4))
    assert (rb[-1]["best_id"], rb[-1]["best_count"]) == ref[:2]
_lib.check(L.lr_debug_pdl(1), "pdl")
print(json.dumps(out, indent=1))
