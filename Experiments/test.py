#!/usr/bin/env python
"""Entry point with the reference's RANSAC flags (Experiments/test.py:294-313 of the reference):

    python -m test --algo RANSAC --mode MMN --iters 1000000 --GC_conf 0.9995 [--codebase GC|open3D]
                   [--fast_rejection ELC|NONE] [--prosac True|False] [--max_samples K]

The per-pair loop keeps the FR(...) call contract of the reference (test.py:163-170) and its 22-column
stats row (test.py:98-100, 197-218).  The balanced-pair data loader + FCGF network of the reference
(MinkowskiEngine, raw datasets) are replaced by the synthetic LiDAR-shaped pair source of SURVEY.md 8(d);
ICP (test.py:183-188) is run through lidarregistration_b200.algorithms.registration_icp.  Multi-GPU: launch with torchrun; pairs are
sharded pair p -> rank p mod G (the reference's test_parallel.sh + DistributedSampler), stats rows are
gathered on rank 0.
"""
import argparse
import logging
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from algorithms.FR import FR  # noqa: E402
from lidarregistration_b200.algorithms import registration_icp  # noqa: E402
from lidarregistration_b200 import metrics, parallel, synthetic  # noqa: E402


def str2bool(v):
    return str(v).lower() in ("true", "1", "yes")


def get_args():
    parser = argparse.ArgumentParser()
    parser.add_argument('--dataset', type=str, default='synthetic', help='only "synthetic" is available here')
    parser.add_argument('--algo', type=str, default='RANSAC', choices=['RANSAC', 'GC'])
    parser.add_argument('--codebase', type=str, default='GC', choices=['open3D', 'GC'])
    parser.add_argument('--mode', type=str, default=None, help='MNN (alias MMN) | GPF | no_filter')
    parser.add_argument('--max_samples', type=int, default=16, help='number of pairs')
    parser.add_argument('--iters', type=int, default=None, help='RANSAC iters')
    parser.add_argument('--spatial_coherence_weight', type=float, default=0.0)
    parser.add_argument('--fast_rejection', type=str, default='ELC', choices=['SPRT', 'ELC', 'NONE'])
    parser.add_argument('--prosac', type=str2bool, default=True)
    parser.add_argument('--GPF_factor', type=float, default=2.0)
    parser.add_argument('--GPF_grid_wid', type=int, default=10)
    parser.add_argument('--GPF_max_matches', type=int, default=10 ** 9)
    parser.add_argument('--GC_conf', type=float, default=0.999)
    parser.add_argument('--GC_LO', type=str2bool, default=True)
    parser.add_argument('--GC_scoring', type=str, default='count', choices=['count', 'MSAC'],
                        help='added by this drop-in: count = the graded criterion (inlier count, lowest id); '
                             'MSAC = pygcransac semantics incl. local optimisation when --GC_LO True')
    parser.add_argument('--num_points', type=int, default=25000, help='points per synthetic scan (cfg 5: ~25k)')
    parser.add_argument('--seed', type=int, default=51)
    return parser.parse_args()


def eval_per_pair(args, pair_ids):
    """-> stats[len(pair_ids), 22] with the reference's column layout (test.py:98-100)."""
    stats = np.zeros([len(pair_ids), 22])
    for k, p in enumerate(pair_ids):
        t0 = time.time()
        rng = np.random.default_rng(args.seed + 5000 + p)
        n = int(args.num_points * rng.uniform(0.8, 1.2))
        d = synthetic.make_pair(n, seed=args.seed + 5000 + p, sigma_f=float(rng.uniform(0.05, 0.12)))
        data_time = time.time() - t0
        T, model_time, src_pcd, tgt_pcd, n_init, ir_init, n_filt, ir_filt = FR(
            torch.from_numpy(d["xyz0"]), torch.from_numpy(d["xyz1"]), torch.from_numpy(d["feat0"]),
            torch.from_numpy(d["feat1"]), args, d["T_gt"])
        re, te = metrics.rotation_error_deg(T, d["T_gt"]), metrics.translation_error_cm(T, d["T_gt"])
        t0 = time.time()
        T_icp = registration_icp(src_pcd, tgt_pcd, 0.6, T).transformation  # test.py:183-188
        torch.cuda.synchronize()
        icp_time = time.time() - t0
        re_icp, te_icp = metrics.rotation_error_deg(T_icp, d["T_gt"]), metrics.translation_error_cm(T_icp, d["T_gt"])
        stats[k, 0] = float(re < 5.0 and te < 60.0)  # test.py:330-331
        stats[k, 1], stats[k, 2] = re, te
        stats[k, 5:9] = np.nan  # pred_labels are NaN for RANSAC (test.py:170)
        stats[k, 9], stats[k, 10], stats[k, 11] = model_time, data_time, icp_time
        stats[k, 12], stats[k, 13], stats[k, 14] = float(re_icp < 5.0 and te_icp < 60.0), re_icp, te_icp
        stats[k, 15:19] = n_init, ir_init, n_filt, ir_filt
        stats[k, 19], stats[k, 20], stats[k, 21] = 0, p, p
    return stats


def analyze_stats(stats):
    ok = stats[:, 0] > 0
    logging.info("pairs %d  recall %.2f%%  RE %.3f deg  TE %.2f cm (means over successes)  model time mean %.4f s, "
                 "99%% quantile %.4f s", len(stats), 100 * ok.mean(), stats[ok, 1].mean() if ok.any() else float("nan"),
                 stats[ok, 2].mean() if ok.any() else float("nan"), stats[:, 9].mean(), np.quantile(stats[:, 9], 0.99))
    logging.info("pairs/s (model time only): %.2f", 1.0 / stats[:, 9].mean())
    ok2 = stats[:, 12] > 0
    logging.info("after ICP: recall %.2f%%  RE %.3f deg  TE %.2f cm  icp time mean %.4f s", 100 * ok2.mean(),
                 stats[ok2, 13].mean() if ok2.any() else float("nan"), stats[ok2, 14].mean() if ok2.any() else float("nan"),
                 stats[:, 11].mean())


def main():
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    args = get_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        torch.distributed.init_process_group("nccl")
    stats = eval_per_pair(args, parallel.shard_pairs(args.max_samples, rank, world))
    stats = parallel.gather_rows(stats)
    if rank == 0:
        analyze_stats(stats)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
