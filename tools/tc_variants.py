"""A/B of tensor-sweep build variants on the GPU box: rebuilds liblidarreg.so with each flag set in a fresh
process, times k_score_tc (library CUDA events) on the cfg-3 pair with ELC on / off, checks the result signature.
usage: python tools/tc_variants.py "-DLR_TCS_TN=32" "-DLR_TCS_TN=64 -DLR_TCS_SPIN=0" ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from lidarregistration_b200 import engine, synthetic
out = {}
for name, elc, iters in (("cfg3_elc", True, 1000000), ("cfg3_noelc_200k", False, 200000)):
    d = synthetic.make_correspondences(30000, 0.3, seed=51 + 3000)
    a, b = engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])
    p = engine.make_params(threshold=0.6, confidence=1.0, max_iters=iters, seed=51, use_elc=elc)
    for _ in range(3):
        r = engine.ransac_rigid(a, b, p)
    engine.prof_read(engine.PROF_SCORE); engine.prof_enable(True)
    reps = 10
    for _ in range(reps):
        r = engine.ransac_rigid(a, b, p)
    engine.prof_enable(False)
    ms, launches = engine.prof_read(engine.PROF_SCORE)
    out[name] = dict(sweep_ms=ms / reps, sig=[r["best_id"], r["best_count"], r["n_scored"], r["n_rechecked"]])
print(json.dumps(out))
''' % ROOT


def main():
    res = {}
    for flags in sys.argv[1:] or [""]:
        env = dict(os.environ, LIDARREG_NVCC_FLAGS=flags)
        b = subprocess.run([sys.executable, "-m", "lidarregistration_b200.build", "--force"], cwd=ROOT, env=env,
                           capture_output=True, text=True)
        if b.returncode != 0:
            res[flags] = {"build_error": b.stderr[-400:]}
            continue
        r = subprocess.run([sys.executable, "-c", CHILD], cwd=ROOT, capture_output=True, text=True, timeout=600)
        try:
            res[flags] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res[flags] = {"error": (r.stdout + r.stderr)[-600:]}
        print(flags, json.dumps(res[flags]), flush=True)
    # leave the product build behind
    subprocess.run([sys.executable, "-m", "lidarregistration_b200.build", "--force"], cwd=ROOT,
                   env={k: v for k, v in os.environ.items() if k != "LIDARREG_NVCC_FLAGS"})
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tc_variants.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
