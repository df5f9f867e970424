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
for mode in (1, 2, 3):  # exact CUDA-core sweep, tcgen05 with fp32 / fp16 accumulators
    engine.match_set_mode(mode)
    i1b, i2b = engine.match_nn(f0, f1, want_2nd=True)
    assert np.array_equal(i1b.cpu().numpy(), o1) and np.array_equal(i2b.cpu().numpy(), o2)
engine.match_set_mode(0)
# unit-norm features: the K = 32 path with its masked tail; monotone targets: the event-overflow scan
g0 = f0 / np.linalg.norm(f0, axis=1, keepdims=True); g1 = f1 / np.linalg.norm(f1, axis=1, keepdims=True)
w = np.linspace(0, 1, 3000, dtype=np.float32)[:, None]
h1 = (1 - w) * rng.standard_normal((3000, 32)).astype(np.float32) + w * g0[0]
h1 /= np.linalg.norm(h1, axis=1, keepdims=True)
for a, b in ((g0, g1), (g0[:40], h1)):
    j1, j2 = engine.match_nn(a, b, want_2nd=True)
    _, p1, p2 = O.find_nn(a, b, True)
    assert np.array_equal(j1.cpu().numpy(), p1) and np.array_equal(j2.cpu().numpy(), p2)
d = synthetic.make_correspondences(3000, 0.3, seed=4)
for sampler, m in ((0, 3), (1, 3), (2, 4)):
    p = engine.make_params(max_iters=20000, round_size=4096, sampler=sampler, sample_size=m, confidence=0.999)
    r = engine.ransac_rigid(d["src"], d["tgt"], p, want_mask=True)
    ref = O.ransac(d["src"], d["tgt"], m=m, sampler=sampler, conf=0.999, max_iters=20000, round_size=4096)
    assert r["best_id"] == ref["best_id"] and r["best_count"] == ref["best_count"]
pairs = [(torch.from_numpy(d["src"]).cuda(), torch.from_numpy(d["tgt"]).cuda())] * 3
pb = engine.make_params(max_iters=20000, confidence=1.0)
rb = engine.ransac_rigid_batch(pairs, pb)
rs = engine.ransac_rigid(*pairs[0], pb)
assert all(x["best_id"] == rs["best_id"] and x["best_count"] == rs["best_count"] for x in rb)
s = rng.integers(0, 3000, (5000, 3)).astype(np.int32)
c, b, _ = engine.ransac_score_samples(d["src"], d["tgt"], s)
oc, ob = O.score_samples(d["src"], d["tgt"], s, 0.6)
assert np.array_equal(c.cpu().numpy(), oc) and b == ob
# GC semantics: MSAC selection, local optimisation, iterated least squares (single call, batch entry, fed hook)
pg = engine.make_params(max_iters=20000, round_size=4096, confidence=0.999, scoring=engine.SCORE_MSAC, lo_rounds=3,
                        lo_trials=8, lsq_iters=2)
rg = engine.ransac_rigid(d["src"], d["tgt"], pg, want_mask=True)
og = O.ransac_gc(d["src"], d["tgt"], conf=0.999, max_iters=20000, round_size=4096, lo_rounds=3, lo_trials=8, lsq_iters=2)
assert (rg["best_id"], rg["best_score"], rg["lo_score"]) == (og["best_id"], og["best_score"], og["lo_score"])
rgb = engine.ransac_rigid_batch(pairs, pg)
assert all(x["best_id"] == rg["best_id"] and x["lo_score"] == rg["lo_score"] for x in rgb)
sc, si, sb = engine.ransac_score_samples_msac(d["src"], d["tgt"], s)
osc, osi, osb = O.score_samples_msac(d["src"], d["tgt"], s, 0.6)
assert np.array_equal(sc.cpu().numpy(), osc) and np.array_equal(si.cpu().numpy(), osi) and sb == osb
# the inlier sweep without the early-out gives the same result
engine.ransac_set_mode(1)
r1 = engine.ransac_rigid(*pairs[0], pb)
engine.ransac_set_mode(0)
assert r1["best_id"] == rs["best_id"] and r1["best_count"] == rs["best_count"]
# f4: hashed-grid ICP (small clouds: the sanitizer slows everything ~50x) and seed scoring on the tensor sweep
from lidarregistration_b200.algorithms import registration_icp
pp = synthetic.make_pair(1500, seed=9, overlap=0.8)
T0 = pp["T_gt"].copy()
T0[:3, 3] += [0.1, -0.1, 0.05]
ri = registration_icp(pp["xyz0"], pp["xyz1"], 0.6, T0, max_iteration=3)
oi = O.icp(pp["xyz0"], pp["xyz1"], 0.6, T0, max_iteration=3)
assert ri.iterations == oi[3] and ri.fitness == oi[1]
models = np.tile(d["T_gt"], (200, 1, 1))
models[:, :3, 3] += rng.normal(0, 0.3, (200, 3))
rsd = engine.seeds_score(d["src"], d["tgt"], models, 0.6)
ocn, obs = O.seeds_score(d["src"], d["tgt"], models, 0.6)
assert np.array_equal(rsd["counts"].cpu().numpy(), ocn) and rsd["best"] == obs
Aw = rng.normal(size=(50, 20, 3)).astype(np.float32)
Tw = engine.kabsch_weighted_batch(Aw, Aw + 1.0, None).cpu().numpy()
assert np.array_equal(Tw[7], O.kabsch_weighted(Aw[7], Aw[7] + 1.0, None))
print("sanitize workload ok")
