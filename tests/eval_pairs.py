#!/usr/bin/env python
"""cfg 5 of BASELINE.json in miniature: a batch of synthetic LiDAR-shaped pairs through the GPU path
(FR, --algo RANSAC --mode MNN) and through the CPU oracle pipeline with the same parameters;
reports RRE / RTE / recall (reference definitions, Experiments/libs/loss.py:44-51) of both.

    python tests/eval_pairs.py --pairs 40 --points 8000 --iters 100000 --out profiles/r1_cfg5_accuracy.json
"""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidarregistration_b200 import metrics, synthetic  # noqa: E402
from lidarregistration_b200.algorithms import FR  # noqa: E402
from oracle import lr_oracle as O  # noqa: E402


def oracle_fr(d, iters, conf, prosac):
    """the reference pipeline of FR.py:16-119 (codebase GC, mode MNN) on the CPU oracle"""
    _, i1, i2 = O.find_nn(d["feat0"], d["feat1"], return_2nd=True)
    m0, m1 = O.nn_to_mutual(d["feat0"], d["feat1"], i1)
    A, B = d["xyz0"][m0], d["xyz1"][m1]
    sampler = O.UNIFORM
    if prosac:
        q = -O.ratio(d["feat0"], d["feat1"], m0, m1, i2[m0])
        order = np.argsort(-q, kind="stable")
        A, B = A[order], B[order]
        sampler = O.PROSAC
    r = O.ransac(A, B, m=3, sampler=sampler, use_elc=True, thr=0.6, conf=conf, max_iters=iters, round_size=65536, seed=51)
    return r["T_refit"] if r["best_count"] > 0 else np.eye(4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=40)
    ap.add_argument("--points", type=int, default=8000)
    ap.add_argument("--iters", type=int, default=100000)
    ap.add_argument("--conf", type=float, default=0.9995)
    ap.add_argument("--prosac", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    args = SimpleNamespace(mode="MMN", iters=a.iters, codebase="GC", prosac=bool(a.prosac), spatial_coherence_weight=0.0,
                           GC_conf=a.conf, fast_rejection="ELC", GC_LO=True, GPF_factor=2.0, GPF_grid_wid=10,
                           GPF_max_matches=10 ** 9, seed=51)
    Tg, To, Tgt, tg, to = [], [], [], 0.0, 0.0
    for p in range(a.pairs):
        rng = np.random.default_rng(51 + 5000 + p)
        n = int(a.points * rng.uniform(0.8, 1.2))
        d = synthetic.make_pair(n, seed=51 + 5000 + p, sigma_f=float(rng.uniform(0.05, 0.14)),
                                overlap=float(rng.uniform(0.15, 0.9)))
        t0 = time.time()
        T = FR(torch.from_numpy(d["xyz0"]), torch.from_numpy(d["xyz1"]), torch.from_numpy(d["feat0"]),
               torch.from_numpy(d["feat1"]), args, d["T_gt"])[0]
        tg += time.time() - t0
        t0 = time.time()
        Tref = oracle_fr(d, a.iters, a.conf, bool(a.prosac))
        to += time.time() - t0
        Tg.append(T), To.append(Tref), Tgt.append(d["T_gt"])
    g, o = metrics.summarize(Tg, Tgt), metrics.summarize(To, Tgt)
    dR = max(float(np.abs(x[:3, :3] - y[:3, :3]).max()) for x, y in zip(Tg, To))
    dt = max(float(np.abs(x[:3, 3] - y[:3, 3]).max()) for x, y in zip(Tg, To))
    out = dict(pairs=a.pairs, points=a.points, iters=a.iters, conf=a.conf, prosac=bool(a.prosac), gpu=g, cpu_oracle=o,
               max_abs_rotation_entry_diff=dR, max_abs_translation_diff_m=dt, gpu_seconds=tg, cpu_oracle_seconds=to,
               cpu_threads=O.num_threads())
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
