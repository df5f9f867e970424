"""ctypes front-end of the CPU oracle (oracle/lr_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never from the product
package.  See the header of lr_oracle.c for what is restated and the parity
status (matching pinned to the reference's matching.py; RANSAC unpinned).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblr_oracle.so")
_lib = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


class LroStats(ctypes.Structure):
    _fields_ = [("iters_run", ctypes.c_int64), ("n_passed", ctypes.c_int64), ("best_id", ctypes.c_int64),
                ("best_count", ctypes.c_int64), ("refit_count", ctypes.c_int64)]


class LroGcStats(ctypes.Structure):
    _fields_ = [("iters_run", ctypes.c_int64), ("n_passed", ctypes.c_int64), ("best_id", ctypes.c_int64),
                ("best_score", ctypes.c_int64), ("best_inliers", ctypes.c_int64), ("lo_score", ctypes.c_int64),
                ("final_score", ctypes.c_int64), ("refit_count", ctypes.c_int64), ("lo_improved", ctypes.c_int32),
                ("lsq_improved", ctypes.c_int32)]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "lr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_REF_SO = os.path.join(_HERE, "_ref", "libelc_ref.so")
_ref = None


def build_ref():
    """oracle/_ref: compile the reference's own ELC header in place (needs /root/reference; a prebuilt file is kept
    otherwise).  Returns the path, or None if neither the reference nor a prebuilt file exists."""
    if os.path.exists("/root/reference/GC-RANSAC/src/pygcransac/include/preemption/preemption_edge_length.h"):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return _REF_SO if os.path.exists(_REF_SO) else None


def ref_elc(src, tgt, sample):
    """The REFERENCE's EdgeLenPreemptiveVerification::verifyModel (preemption_edge_length.h:71-128, compiled from
    the reference tree into oracle/_ref) on the N x 6 fp64 matrix pygcransac builds (gcransac_python.cpp:426-435)."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            raise FileNotFoundError(_REF_SO)
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_elc_verify.restype = ctypes.c_int
    pts = np.ascontiguousarray(np.concatenate([np.asarray(src, np.float32).astype(np.float64),
                                               np.asarray(tgt, np.float32).astype(np.float64)], axis=1))
    smp = np.ascontiguousarray(sample, dtype=np.uint64)
    return bool(_ref.ref_elc_verify(_p(pts, c_f64p), ctypes.c_long(len(pts)),
                                    smp.ctypes.data_as(ctypes.POINTER(ctypes.c_size_t)), ctypes.c_size_t(len(smp))))


def has_ref():
    return os.path.exists(_REF_SO)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.lro_count_inliers.restype = ctypes.c_int64
        _lib.lro_mutual.restype = ctypes.c_int64
        _lib.lro_score_samples.restype = ctypes.c_int64
        _lib.lro_conf_iters.restype = ctypes.c_int64
        _lib.lro_refit_indexed.restype = ctypes.c_int64
        _lib.lro_msac.restype = ctypes.c_double
        _lib.lro_msac_q.restype = ctypes.c_int64
        _lib.lro_score_samples_msac.restype = ctypes.c_int64
        _lib.lro_elc.restype = ctypes.c_int
        _lib.lro_num_threads.restype = ctypes.c_int
        _lib.lro_seeds_score.restype = ctypes.c_int64
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


def num_threads():
    return int(lib().lro_num_threads())


def set_threads(t):
    lib().lro_set_threads(int(t))


def set_o3d_faithful(on):
    """CPU-timing variant of the loop: copy + transform the whole source cloud per surviving hypothesis as Open3D's
    RegistrationRANSACBasedOnCorrespondence does; counts are unchanged"""
    lib().lro_set_o3d_faithful(int(bool(on)))


def sqnorms(F):
    F = _f32(F)
    out = np.empty(F.shape[0], np.float32)
    lib().lro_sqnorms(_p(F, c_f32p), ctypes.c_int64(F.shape[0]), F.shape[1], _p(out, c_f32p))
    return out


def find_nn(F0, F1, return_2nd=False, return_dist=False):
    """matching.py:22-65 -> (idx0, idx1, idx1_2nd|None[, d1, d2])"""
    F0, F1 = _f32(F0), _f32(F1)
    N, M, D = F0.shape[0], F1.shape[0], F0.shape[1]
    idx1 = np.empty(N, np.int64)
    idx2 = np.empty(N, np.int64) if return_2nd else None
    d1 = np.empty(N, np.float32) if return_dist else None
    d2 = np.empty(N, np.float32) if return_dist else None
    lib().lro_find_nn(_p(F0, c_f32p), ctypes.c_int64(N), _p(F1, c_f32p), ctypes.c_int64(M), D, _p(idx1, c_i64p),
                      _p(idx2, c_i64p) if return_2nd else None, _p(d1, c_f32p) if return_dist else None,
                      _p(d2, c_f32p) if return_dist else None)
    out = (np.arange(N, dtype=np.int64), idx1, idx2)
    return out + (d1, d2) if return_dist else out


def nn_to_mutual(F0, F1, idx1):
    """matching.py:222-239 -> (idx0', idx1') sorted by idx0"""
    F0, F1 = _f32(F0), _f32(F1)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
    N, M, D = F0.shape[0], F1.shape[0], F0.shape[1]
    oi = np.empty(N, np.int64)
    oj = np.empty(N, np.int64)
    K = lib().lro_mutual(_p(F0, c_f32p), ctypes.c_int64(N), _p(F1, c_f32p), ctypes.c_int64(M), D, _p(idx1, c_i64p),
                         _p(oi, c_i64p), _p(oj, c_i64p))
    return oi[:K].copy(), oj[:K].copy()


def ratio(F0, F1, i0, i1, i2):
    """matching.py:89-98"""
    F0, F1 = _f32(F0), _f32(F1)
    i0, i1, i2 = (np.ascontiguousarray(a, dtype=np.int64) for a in (i0, i1, i2))
    out = np.empty(len(i0), np.float32)
    lib().lro_ratio(_p(F0, c_f32p), _p(F1, c_f32p), F0.shape[1], ctypes.c_int64(len(i0)), _p(i0, c_i64p),
                    _p(i1, c_i64p), _p(i2, c_i64p), _p(out, c_f32p))
    return out


UNIFORM, PROSAC, REPLACE = 0, 1, 2  # sampler ids of include/lidarreg.h


def prosac_growth(n, m):
    g = np.empty(n, np.uint32)
    lib().lro_prosac_growth(ctypes.c_int64(n), int(m), g.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return g


def sample(seed, hid, sampler, m, n, growth=None):
    out = np.empty(m, np.int32)
    if sampler == PROSAC and growth is None:
        growth = prosac_growth(n, m)
    gp = growth.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)) if growth is not None else None
    lib().lro_sample(ctypes.c_uint64(seed), ctypes.c_uint64(hid), int(sampler), int(m), ctypes.c_int64(n), gp,
                     _p(out, c_i32p))
    return out


def elc(P, Q, ratio_=0.9):
    P = np.ascontiguousarray(P, dtype=np.float64)
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    return bool(lib().lro_elc(_p(P, c_f64p), _p(Q, c_f64p), P.shape[0], ctypes.c_double(ratio_)))


def _T44(T12):
    T = np.eye(4)
    T[:3, :] = np.asarray(T12).reshape(3, 4)
    return T


def kabsch(P, Q):
    """Least-squares rigid motion Q ~ R P + t -> 4x4 (column-vector convention)."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    T = np.empty(12, np.float64)
    lib().lro_kabsch(_p(P, c_f64p), _p(Q, c_f64p), ctypes.c_int64(P.shape[0]), _p(T, c_f64p))
    return _T44(T)


def count_inliers(src, tgt, T, thr, return_mask=False):
    src, tgt = _f32(src), _f32(tgt)
    T12 = np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:3, :].reshape(-1))
    mask = np.empty(src.shape[0], np.uint8) if return_mask else None
    c = lib().lro_count_inliers(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(src.shape[0]), _p(T12, c_f64p),
                                ctypes.c_double(thr), _p(mask, c_u8p) if return_mask else None)
    return (int(c), mask.astype(bool)) if return_mask else int(c)


def msac(src, tgt, T, thr):
    src, tgt = _f32(src), _f32(tgt)
    T12 = np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:3, :].reshape(-1))
    cnt = ctypes.c_int64(0)
    v = lib().lro_msac(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(src.shape[0]), _p(T12, c_f64p),
                       ctypes.c_double(thr), ctypes.byref(cnt))
    return float(v), int(cnt.value)


def score_samples(src, tgt, samples, thr, use_elc=True, elc_ratio=0.9, return_models=False):
    """Fed-sample hook: counts[h] (-1 = ELC reject), selected h (max count, lowest h)."""
    src, tgt = _f32(src), _f32(tgt)
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    H, m = samples.shape
    counts = np.empty(H, np.int32)
    models = np.empty((H, 12), np.float64) if return_models else None
    best = lib().lro_score_samples(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(src.shape[0]),
                                   _p(samples, c_i32p), ctypes.c_int64(H), m, ctypes.c_double(thr), int(use_elc),
                                   ctypes.c_double(elc_ratio), _p(counts, c_i32p),
                                   _p(models, c_f64p) if return_models else None)
    return (counts, int(best), models) if return_models else (counts, int(best))


def conf_iters(c, n, m, conf, max_iters):
    return int(lib().lro_conf_iters(ctypes.c_int64(c), ctypes.c_int64(n), int(m), ctypes.c_double(conf),
                                    ctypes.c_int64(max_iters)))


def ransac(src, tgt, m=3, sampler=0, use_elc=True, elc_ratio=0.9, thr=0.6, conf=1.0, max_iters=100000,
           round_size=65536, seed=51, refit=True, return_mask=False):
    """Full loop, count scoring.  -> dict(T, T_refit, mask, stats...)"""
    src, tgt = _f32(src), _f32(tgt)
    n = src.shape[0]
    T = np.empty(12, np.float64)
    Tr = np.empty(12, np.float64) if refit else None
    mask = np.empty(n, np.uint8) if return_mask else None
    st = LroStats()
    lib().lro_ransac(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(n), int(m), int(sampler), int(use_elc),
                     ctypes.c_double(elc_ratio), ctypes.c_double(thr), ctypes.c_double(conf),
                     ctypes.c_int64(max_iters), ctypes.c_int64(round_size), ctypes.c_uint64(seed), _p(T, c_f64p),
                     _p(Tr, c_f64p) if refit else None, _p(mask, c_u8p) if return_mask else None, ctypes.byref(st))
    return dict(T=_T44(T), T_refit=_T44(Tr) if refit else None, mask=mask.astype(bool) if return_mask else None,
                iters_run=st.iters_run, n_passed=st.n_passed, best_id=st.best_id, best_count=st.best_count,
                refit_count=st.refit_count)


def msac_q(src, tgt, T, thr):
    """Quantised MSAC score (integer sum of trunc((1 - r^2/tau^2) * 65536), tau = 1.5 thr) -> (q, #r^2<tau^2)"""
    src, tgt = _f32(src), _f32(tgt)
    T12 = np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:3, :].reshape(-1))
    cnt = ctypes.c_int64(0)
    q = lib().lro_msac_q(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(src.shape[0]), _p(T12, c_f64p),
                         ctypes.c_double(thr), ctypes.byref(cnt))
    return int(q), int(cnt.value)


def score_samples_msac(src, tgt, samples, thr, use_elc=True, elc_ratio=0.9):
    """Fed-sample hook, MSAC flavour -> (scores[H] int64 (-1 = ELC reject), inliers[H] int32, selected h)"""
    src, tgt = _f32(src), _f32(tgt)
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    H, m = samples.shape
    scores = np.empty(H, np.int64)
    inl = np.empty(H, np.int32)
    best = lib().lro_score_samples_msac(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(src.shape[0]),
                                        _p(samples, c_i32p), ctypes.c_int64(H), m, ctypes.c_double(thr),
                                        int(use_elc), ctypes.c_double(elc_ratio), _p(scores, c_i64p),
                                        _p(inl, c_i32p))
    return scores, inl, int(best)


def ransac_gc(src, tgt, m=3, sampler=0, use_elc=True, elc_ratio=0.9, thr=0.6, conf=1.0, max_iters=100000,
              round_size=65536, seed=51, lo_rounds=10, lo_trials=20, lsq_iters=10, refit=True, return_mask=False):
    """GC-RANSAC semantics: MSAC selection, local optimisation, iterated least squares."""
    src, tgt = _f32(src), _f32(tgt)
    n = src.shape[0]
    T = np.empty(12, np.float64)
    Tr = np.empty(12, np.float64) if refit else None
    mask = np.empty(n, np.uint8) if return_mask else None
    st = LroGcStats()
    lib().lro_ransac_gc(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(n), int(m), int(sampler), int(use_elc),
                        ctypes.c_double(elc_ratio), ctypes.c_double(thr), ctypes.c_double(conf),
                        ctypes.c_int64(max_iters), ctypes.c_int64(round_size), ctypes.c_uint64(seed),
                        int(lo_rounds), int(lo_trials), int(lsq_iters), _p(T, c_f64p),
                        _p(Tr, c_f64p) if refit else None, _p(mask, c_u8p) if return_mask else None,
                        ctypes.byref(st))
    out = dict(T=_T44(T), T_refit=_T44(Tr) if refit else None, mask=mask.astype(bool) if return_mask else None)
    out.update({k: int(getattr(st, k)) for k, _ in st._fields_})
    return out


def refit_indexed(xyz0, xyz1, i0, i1, T, thr):
    """FR.py:99-111: inliers of an indexed correspondence set under T, then Kabsch."""
    xyz0, xyz1 = _f32(xyz0), _f32(xyz1)
    i0 = np.ascontiguousarray(i0, dtype=np.int64)
    i1 = np.ascontiguousarray(i1, dtype=np.int64)
    T12 = np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:3, :].reshape(-1))
    out = np.empty(12, np.float64)
    k = lib().lro_refit_indexed(_p(xyz0, c_f32p), _p(xyz1, c_f32p), _p(i0, c_i64p), _p(i1, c_i64p),
                                ctypes.c_int64(len(i0)), _p(T12, c_f64p), ctypes.c_double(thr), _p(out, c_f64p))
    return _T44(out), int(k)


def kabsch_weighted(A, B, w=None):
    """Experiments/models/common.py:7-45 (rigid_transform_3d) for one neighbourhood: A, B [k,3], w [k] -> 4x4."""
    A, B = _f32(A), _f32(B)
    w = None if w is None else _f32(w)
    T = np.empty(12, np.float64)
    lib().lro_kabsch_weighted(_p(A, c_f32p), _p(B, c_f32p), _p(w, c_f32p) if w is not None else None,
                              ctypes.c_int64(A.shape[0]), _p(T, c_f64p))
    return _T44(T)


def seeds_score(src, tgt, models, thr, return_labels=False):
    """Experiments/models/PointDSC.py:319-336: models [S,4,4] -> (counts[S], best seed[, labels of the best])."""
    src, tgt = _f32(src), _f32(tgt)
    M = np.ascontiguousarray(np.asarray(models, dtype=np.float64)[:, :3, :].reshape(len(models), 12))
    counts = np.empty(len(M), np.int32)
    labels = np.empty(src.shape[0], np.uint8) if return_labels else None
    best = lib().lro_seeds_score(_p(src, c_f32p), _p(tgt, c_f32p), ctypes.c_int64(src.shape[0]), _p(M, c_f64p),
                                 ctypes.c_int64(len(M)), ctypes.c_double(thr), _p(counts, c_i32p),
                                 _p(labels, c_u8p) if return_labels else None)
    return (counts, int(best), labels.astype(bool)) if return_labels else (counts, int(best))


def nn3d_radius(src, tgt, T, radius):
    """nearest target of every T-transformed source point with d^2 < radius^2 (brute force) -> (idx, d2); -1 = none"""
    src, tgt = _f32(src), _f32(tgt)
    T12 = np.ascontiguousarray(np.asarray(T, dtype=np.float64)[:3, :].reshape(-1))
    idx = np.empty(src.shape[0], np.int64)
    d2 = np.empty(src.shape[0], np.float64)
    lib().lro_nn3d_radius(_p(src, c_f32p), ctypes.c_int64(src.shape[0]), _p(tgt, c_f32p), ctypes.c_int64(tgt.shape[0]),
                          _p(T12, c_f64p), ctypes.c_double(radius), _p(idx, c_i64p), _p(d2, c_f64p))
    return idx, d2


def icp(src, tgt, max_dist, init=None, max_iteration=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """Point-to-point ICP (Experiments/test.py:183-188 / Open3D registration_icp semantics) from the oracle's
    own primitives: nearest target of every transformed source point inside max_dist (brute force, canonical
    fp64 squared distance, ties -> lowest index), Kabsch over those pairs, Open3D's relative fitness / rmse
    stopping rule.  -> (T, fitness, inlier_rmse, iterations)"""
    src, tgt = _f32(src), _f32(tgt)
    T = np.eye(4) if init is None else np.asarray(init, dtype=np.float64).copy()
    n = len(src)

    def evaluate(T):
        idx, d2 = nn3d_radius(src, tgt, T, max_dist)
        keep = np.nonzero(idx >= 0)[0]
        cnt = len(keep)
        T_new, k = refit_indexed(src, tgt, keep, idx[keep], T, max_dist)
        assert k == cnt
        return T_new, cnt / n, float(np.sqrt(d2[keep].sum() / cnt)) if cnt else 0.0

    T_next, fitness, rmse = evaluate(T)
    it = 0
    for it in range(1, max_iteration + 1):
        T = T_next
        T_next, f2, r2 = evaluate(T)
        done = abs(fitness - f2) < rel_fitness and abs(rmse - r2) < rel_rmse
        fitness, rmse = f2, r2
        if done:
            break
    return T, fitness, rmse, it
