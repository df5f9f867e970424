"""Drop-in for Experiments/algorithms/matching.py of AmnonDrory/LidarRegistration.

Same function names, arguments, return types (int64 CPU index tensors, pairs
sorted by idx0, lowest index on ties) and error behaviour as the reference
file (citations: /root/reference/Experiments/algorithms/matching.py); the
arithmetic runs in liblidarreg.so on the GPU.  The `*_dev` variants keep
everything in HBM for the fused FR() path.
"""
from copy import deepcopy
from time import time

import numpy as np
import torch

from .. import engine


def _sync():
    torch.cuda.synchronize()


def find_nn(F0, F1, return_2nd=False):
    """matching.py:22-65 -> (corres_idx0 = arange(N), corres_idx1, idx1_2nd | None), int64 CPU."""
    idx1, idx2 = engine.match_nn(F0, F1, want_2nd=return_2nd)
    N = idx1.shape[0]
    corres_idx0 = torch.arange(N).long().squeeze()
    corres_idx1 = idx1.long().squeeze().cpu()
    if return_2nd:
        return corres_idx0, corres_idx1, idx2.long().squeeze().cpu()
    return corres_idx0, corres_idx1, None


_EXTRA_2ND = {}  # (N, M, D) -> seconds the 2nd neighbour adds to one sweep, measured once per shape


def find_2nn_dev(fcgf_feats0, fcgf_feats1):
    """find_2nn with everything left in HBM: (idx1, idx1_2nd) int64 CUDA tensors + the extra seconds.

    The reference runs find_nn twice and charges `model_time` the difference (matching.py:6-19, FR.py:117).
    Here ONE sweep delivers both neighbours; what the second one adds to it (the WANT2 variant of the sweep and
    of the re-rank against the 1-NN variant) is measured on the device the first time a shape is seen -- two
    extra sweeps, once -- and reused, so the steady state costs a single sweep and no host round trip."""
    f0, f1 = engine.to_dev_f32(fcgf_feats0), engine.to_dev_f32(fcgf_feats1)
    key = (int(f0.shape[0]), int(f1.shape[0]), int(f0.shape[1]))
    if key not in _EXTRA_2ND:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        engine.match_nn(f0, f1, want_2nd=True)  # warm-up (scratch allocation)
        ev[0].record()
        engine.match_nn(f0, f1, want_2nd=False)
        ev[1].record()
        engine.match_nn(f0, f1, want_2nd=True)
        ev[2].record()
        ev[2].synchronize()
        _EXTRA_2ND[key] = max(0.0, (ev[1].elapsed_time(ev[2]) - ev[0].elapsed_time(ev[1])) * 1e-3)
    idx1, idx2 = engine.match_nn(f0, f1, want_2nd=True)
    return idx1, idx2, _EXTRA_2ND[key]


def find_2nn(fcgf_feats0, fcgf_feats1):
    """matching.py:6-19: NN + 2nd NN (int64 CPU tensors) and the *extra* seconds the 2nd NN costs
    (see find_2nn_dev for how that figure is obtained without running the sweep twice per call)."""
    idx1, idx2, extra = find_2nn_dev(fcgf_feats0, fcgf_feats1)
    N = idx1.shape[0]
    return torch.arange(N).long().squeeze(), idx1.cpu(), idx2.cpu(), extra


def nn_to_mutual_dev(feats0, feats1, idx1, idx1_2nd=None):
    """nn_to_mutual on device tensors: (idx0'[K], idx1'[K], idx1_2nd'[K] | None), sorted by idx0; one 8-byte
    read-back (K) is the only host round trip."""
    out_i, out_j = engine.match_mutual(feats0, feats1, idx1)
    return out_i, out_j, (idx1_2nd[out_i] if idx1_2nd is not None else None)


def measure_inlier_ratio_dev(idx0, idx1, xyz0_d, xyz1_d, T_gt, voxel_size):
    """measure_inlier_ratio (matching.py:241-249) on device tensors, fp64 like the reference's numpy."""
    if idx0.shape[0] == 0:
        return float("nan")  # 0 / 0 in the reference
    T = torch.as_tensor(np.asarray(T_gt, dtype=np.float64), device=xyz0_d.device)
    p = xyz0_d[idx0].double() @ T[:3, :3].T + T[:3, 3]
    d2 = ((p - xyz1_d[idx1].double()) ** 2).sum(dim=1)
    return float((d2 < (2 * voxel_size) ** 2).double().mean().item())


def torch_intersect(Na, Nb, i_ab, j_ab, i_ba, j_ba):
    """matching.py:67-87: edges present in both lists, sorted by (i, j)."""
    dev = i_ab.device
    ka = i_ab.long() * Nb + j_ab.long()
    kb = i_ba.long().to(dev) * Nb + j_ba.long().to(dev)
    ka_u, kb_u = torch.unique(ka), torch.unique(kb)
    both = ka_u[torch.isin(ka_u, kb_u)]
    return both // Nb, both % Nb


def nn_to_mutual(feats0, feats1, corres_idx0, corres_idx1, idx1_2nd=None, force_return_2nd=False):
    """matching.py:222-239: keep (i, j) iff j = NN_1(i) and i = NN_0(j); sorted by i."""
    assert len(corres_idx0) == len(feats0), "nn_to_mutual relies on corres_idx0 being the full range (matching.py:234)"
    out_i, out_j = engine.match_mutual(feats0, feats1, corres_idx1)
    final_corres_idx0, final_corres_idx1 = out_i.cpu(), out_j.cpu()
    if idx1_2nd is not None:
        idx1_2nd = idx1_2nd[final_corres_idx0]
        return final_corres_idx0, final_corres_idx1, idx1_2nd
    elif force_return_2nd:
        return final_corres_idx0, final_corres_idx1, None
    else:
        return final_corres_idx0, final_corres_idx1


def calc_distance_ratio_in_feature_space(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd):
    """matching.py:89-98: d(f0_i, f1_nn) / (d(f0_i, f1_2nd) + 1e-6), fp32 (on the features' device)."""
    out = engine.match_ratio(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd)
    return out if getattr(fcgf_feats0, "is_cuda", False) else out.cpu()


def mark_best_buddies(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1):
    """matching.py:207-220: which NN pairs are mutual."""
    bb_idx0, bb_idx1 = nn_to_mutual(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1)
    corres_idx0_np = corres_idx0.detach().cpu().numpy()
    corres_idx1_np = corres_idx1.detach().cpu().numpy()
    P = 1 + np.max(corres_idx0_np)
    bb_idx_flat = P * bb_idx1.numpy() + bb_idx0.numpy()
    corres_idx_flat = P * corres_idx1_np + corres_idx0_np
    is_bb = np.isin(corres_idx_flat, bb_idx_flat)
    return is_bb, is_bb.sum()


def Grid_Prioritized_Filter_dev(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd, xyz0, args, BB_first=False):
    """Grid_Prioritized_Filter with everything in HBM (device tensors in and out): best-buddy marking from the mutual
    sweep, the ratio quality, then lr_gpf_filter (cells, water-filling, per-cell selection by rank counting).
    Returns what the reference returns (matching.py:205)."""
    f0, f1 = engine.to_dev_f32(fcgf_feats0), engine.to_dev_f32(fcgf_feats1)
    i0, i1 = engine.to_dev_i64(corres_idx0), engine.to_dev_i64(corres_idx1)
    i2 = engine.to_dev_i64(idx1_2nd)
    orig = (i0, i1, i2)
    assert len(i0) == len(f0), "GPF starts from the full nearest-neighbour set (matching.py:207-220 relies on it)"
    bb_i, bb_j = engine.match_mutual(f0, f1, i1)
    if BB_first:
        TOTAL_NUM = args.GPF_max_matches
        i0, i1, i2 = bb_i, bb_j, i2[bb_i]
        if TOTAL_NUM >= i0.shape[0]:
            return i0, i1, i2, orig[0], orig[1], orig[2], None
        is_bb = None
    else:
        is_bb = torch.zeros(i0.shape[0], dtype=torch.uint8, device=i0.device)
        is_bb[bb_i] = 1
        TOTAL_NUM = args.GPF_factor * int(bb_i.shape[0])
    ratio = engine.match_ratio(f0, f1, i0, i1, i2)
    keep, norm = engine.gpf_filter(ratio, is_bb, xyz0, i0, args.GPF_grid_wid, TOTAL_NUM)
    sel = torch.nonzero(keep).squeeze(1)  # the one size read-back of the filter (as nn_to_mutual has)
    return i0[sel], i1[sel], i2[sel], orig[0], orig[1], orig[2], norm[sel]


def Grid_Prioritized_Filter(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd, xyz0, args,
                            BB_first=False):
    """matching.py:100-205 (--mode GPF): best-buddy marking, 10x10 xy grid, water-filling quota, per-cell selection by
    normalised ratio -- on the device (csrc/lr_gpf.cu); int64 / fp32 CPU tensors out like the reference."""
    out = Grid_Prioritized_Filter_dev(fcgf_feats0, fcgf_feats1, corres_idx0, corres_idx1, idx1_2nd, xyz0, args, BB_first)
    return tuple(None if t is None else t.cpu() for t in out)


def measure_inlier_ratio(corres_idx0, corres_idx1, pcd0, pcd1, T_gt, voxel_size):
    """matching.py:241-249 (ground-truth statistic only; numpy fp64 like the reference)."""
    corres_idx0_ = corres_idx0.detach().cpu().numpy()
    corres_idx1_ = corres_idx1.detach().cpu().numpy()
    pcd0_trans = deepcopy(pcd0)
    pcd0_trans.transform(T_gt)
    dist2 = np.sum((np.array(pcd0_trans.points)[corres_idx0_, :] - np.array(pcd1.points)[corres_idx1_, :]) ** 2,
                   axis=1)
    is_close = dist2 < (2 * voxel_size) ** 2
    return float(is_close.sum()) / len(is_close)
