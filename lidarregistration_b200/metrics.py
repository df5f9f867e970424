"""Registration metrics with the reference's definitions.

Experiments/libs/loss.py:44-51 (TransformationLoss.forward): RE = acos(clamp((tr(R^T R_gt) - 1) / 2)) in
degrees, TE = |t - t_gt| in centimetres, success iff RE < re_thre and TE < te_thre; the RANSAC path uses
re_thre = 5 deg, te_thre = 60 cm (Experiments/test.py:325-331).
"""
import numpy as np


def rotation_error_deg(T, T_gt):
    R, Rg = np.asarray(T)[:3, :3], np.asarray(T_gt)[:3, :3]
    c = (np.trace(R.T @ Rg) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


def translation_error_cm(T, T_gt):
    return float(np.linalg.norm(np.asarray(T)[:3, 3] - np.asarray(T_gt)[:3, 3]) * 100.0)


def registration_success(T, T_gt, re_thre=5.0, te_thre=60.0):
    return rotation_error_deg(T, T_gt) < re_thre and translation_error_cm(T, T_gt) < te_thre


def summarize(Ts, T_gts, re_thre=5.0, te_thre=60.0):
    """-> dict(recall, RRE (deg, mean over successes), RTE (cm, mean over successes))"""
    re = np.array([rotation_error_deg(a, b) for a, b in zip(Ts, T_gts)])
    te = np.array([translation_error_cm(a, b) for a, b in zip(Ts, T_gts)])
    ok = (re < re_thre) & (te < te_thre)
    return dict(recall=float(ok.mean()) if len(ok) else 0.0, RRE=float(re[ok].mean()) if ok.any() else float("nan"),
                RTE=float(te[ok].mean()) if ok.any() else float("nan"), n=int(len(ok)))
