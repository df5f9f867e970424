"""A/B of prebuilt library variants (tools/build_variants.py) on the GPU box: every lib_<name>.so under
lidarregistration_b200/csrc/variants/ runs in its own process (LIDARREG_SO) on the cfg-3 pair, ELC off (every hypothesis
scored: the sweep is 96 % of the pair) and ELC on; prints sweep / gen times (library CUDA events), the wall time of the call
and the result signature (must be the same for all).  usage: python tools/variant_ab.py [name ...]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, time, torch
sys.path.insert(0, %r)
from lidarregistration_b200 import engine, synthetic
d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
a, b = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
out = {}
for name, elc, reps in (("elc_off", False, 6), ("elc_on", True, 40)):
    p = engine.make_params(threshold=0.6, confidence=1.0, max_iters=1000000, seed=51, use_elc=elc)
    for _ in range(3):
        r = engine.ransac_rigid(a, b, p)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        r = engine.ransac_rigid(a, b, p)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / reps * 1e3
    for k in (engine.PROF_SCORE, engine.PROF_GEN, engine.PROF_FIN): engine.prof_read(k)
    engine.prof_enable(True)
    for _ in range(reps):
        r = engine.ransac_rigid(a, b, p)
    engine.prof_enable(False)
    out[name] = dict(sweep_ms=round(engine.prof_read(engine.PROF_SCORE)[0] / reps, 4), gen_ms=round(engine.prof_read(engine.PROF_GEN)[0] / reps, 4),
                     fin_ms=round(engine.prof_read(engine.PROF_FIN)[0] / reps, 4), wall_ms=round(wall, 4),
                     sig=[r["best_id"], r["best_count"], r["n_scored"], r["n_rechecked"], float(r["T_refit"][0, 3])])
print(json.dumps(out))
''' % ROOT

names = sys.argv[1:] or sorted(os.path.basename(p)[4:-3] for p in glob.glob(os.path.join(ROOT, "lidarregistration_b200/csrc/variants/lib_*.so")))
res = {}
for n in names:
    so = os.path.join(ROOT, "lidarregistration_b200/csrc/variants/lib_%s.so" % n)
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=ROOT, env=dict(os.environ, LIDARREG_SO=so), capture_output=True, text=True,
                       timeout=45)
    try:
        res[n] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[n] = {"error": (r.stdout + r.stderr)[-600:]}
    print(n, json.dumps(res[n]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variant_ab.json"), "w"), indent=1)
