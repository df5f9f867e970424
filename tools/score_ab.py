"""A/B of the inlier sweep variants (lr_ransac_set_mode) on cfg-3 / cfg-4 shaped pairs: device time of k_score
per pair, identical results required.  Usage: python tools/score_ab.py [n_pairs]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lidarregistration_b200 import engine, synthetic  # noqa: E402


def run(pairs, params, reps=5):
    for a, b in pairs[:2]:
        engine.ransac_rigid(a, b, params)
    engine.prof_read(engine.PROF_SCORE)
    engine.prof_enable(True)
    out = []
    for _ in range(reps):
        out = [engine.ransac_rigid(a, b, params) for a, b in pairs]
    engine.prof_enable(False)
    ms, launches = engine.prof_read(engine.PROF_SCORE)
    gen_ms, _ = engine.prof_read(engine.PROF_GEN)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        for a, b in pairs:
            engine.ransac_rigid(a, b, params)
    e1.record()
    torch.cuda.synchronize()
    return ms / (reps * len(pairs)), out, gen_ms / (reps * len(pairs)), e0.elapsed_time(e1) / (reps * len(pairs))


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 0]
    res = {}
    for name, ratio, elc in (("cfg3_elc", 0.3, True), ("cfg3_noelc_100k", 0.3, False),
                             ("inl10_elc", 0.1, True), ("inl60_elc", 0.6, True)):
        iters = 100000 if not elc else 1000000
        pairs = []
        for p in range(k):
            d = synthetic.make_correspondences(30000, ratio, seed=51 + 3000 + p)
            pairs.append((engine.to_dev_f32(d["src"]), engine.to_dev_f32(d["tgt"])))
        params = engine.make_params(threshold=0.6, confidence=1.0, max_iters=iters, seed=51, use_elc=elc)
        row = {}
        ref = None
        for mode in modes:
            engine.ransac_set_mode(mode)
            ms, out, gen_ms, pair_ms = run(pairs, params)
            sig = [(o["best_id"], o["best_count"], o["n_scored"]) for o in out]
            row["mode%d_rechecked" % mode] = out[0]["n_rechecked"]
            row["mode%d_gen_ms" % mode] = gen_ms
            row["mode%d_pair_ms" % mode] = pair_ms
            if ref is None:
                ref = sig
            row["mode%d_ms" % mode] = ms
            row["same"] = row.get("same", True) and sig == ref
            row["n_scored"] = out[0]["n_scored"]
        row["speedup_vs_first"] = {m: row["mode%d_ms" % modes[0]] / row["mode%d_ms" % m] for m in modes[1:]}
        res[name] = row
    engine.ransac_set_mode(0)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
