import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidarregistration_b200 import engine, synthetic
from oracle import lr_oracle as O
rng = np.random.default_rng(3)
for (N, M) in [(300, 700), (1000, 130)]:
    f0 = rng.standard_normal((N, 32)).astype(np.float32); f1 = rng.standard_normal((M, 32)).astype(np.float32)
    i1, i2 = engine.match_nn(f0, f1, want_2nd=True)
    mi, mj = engine.match_mutual(f0, f1, i1)
    _, o1, o2 = O.find_nn(f0, f1, True)
    assert np.array_equal(i1.cpu().numpy(), o1) and np.array_equal(i2.cpu().numpy(), o2)
engine.match_set_mode(1)
i1b, _ = engine.match_nn(f0, f1)
engine.match_set_mode(0)
assert np.array_equal(i1b.cpu().numpy(), o1)
d = synthetic.make_correspondences(3000, 0.3, seed=4)
for sampler, m in ((0, 3), (1, 3), (2, 4)):
    p = engine.make_params(max_iters=20000, round_size=4096, sampler=sampler, sample_size=m, confidence=0.999)
    r = engine.ransac_rigid(d["src"], d["tgt"], p, want_mask=True)
    ref = O.ransac(d["src"], d["tgt"], m=m, sampler=sampler, conf=0.999, max_iters=20000, round_size=4096)
    assert r["best_id"] == ref["best_id"] and r["best_count"] == ref["best_count"]
s = rng.integers(0, 3000, (5000, 3)).astype(np.int32)
c, b, _ = engine.ransac_score_samples(d["src"], d["tgt"], s)
oc, ob = O.score_samples(d["src"], d["tgt"], s, 0.6)
assert np.array_equal(c.cpu().numpy(), oc) and b == ob
print("sanitize workload ok")
