"""PointDSC's seed scoring: the second half of `PointDSC.cal_seed_trans`
(Experiments/models/PointDSC.py:293-336), the consumer that reuses the path's batched Kabsch and inlier sweep.

  rigid_transform_3d(src_knn, tgt_knn, total_weight)          :318   -> seedwise_transforms   (lr_kabsch_weighted_batch)
  pred = R src + t ; L2 < inlier_threshold ; mean ; argmax    :321-326 -> score_seeds          (lr_seeds_score: the
                                                                        tensor-core sweep fed with the seeds' transforms)
  final_trans / final_labels                                  :329-331

Shapes follow the reference (a leading batch dimension of 1 is accepted and returned); tensors come back on the
device the keypoints live on when they are CUDA tensors, else on the CPU like the reference's.
"""
import numpy as np
import torch

from .. import engine


def _squeeze(x, nd):
    x = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
    batched = x.dim() == nd + 1
    if batched:
        assert x.shape[0] == 1, "one point-cloud pair per call (the reference's test loop runs with batch size 1)"
        x = x[0]
    return x, batched


def seedwise_transforms(src_knn, tgt_knn, total_weight=None):
    """[S,k,3], [S,k,3], [S,k] | None -> [S,4,4] float32 (models/common.py:7-45 semantics, computed in fp64)"""
    T = engine.kabsch_weighted_batch(src_knn, tgt_knn, total_weight)
    return T.to(torch.float32)


def score_seeds(seedwise_trans, src_keypts, tgt_keypts, inlier_threshold):
    """-> (seedwise_fitness [S], final_trans [4,4], final_labels [n] float, best seed) as PointDSC.py:319-331;
    a leading batch dimension of 1 is kept when the inputs carry one."""
    trans, batched = _squeeze(seedwise_trans, 3)
    src, _ = _squeeze(src_keypts, 2)
    tgt, _ = _squeeze(tgt_keypts, 2)
    out_dev = src.device if torch.is_tensor(src_keypts) else torch.device("cpu")
    n = int(src.shape[0])
    res = engine.seeds_score(src, tgt, trans, float(inlier_threshold))
    fitness = (res["counts"].to(torch.float32) / float(n)).to(out_dev)
    final_trans = trans[res["best"]].to(out_dev).to(torch.float32)
    labels = res["labels"].to(torch.float32).to(out_dev)
    if batched:
        return fitness[None], final_trans[None], labels[None], res["best"]
    return fitness, final_trans, labels, res["best"]
