"""Device-resident front-end of the C ABI: CUDA tensors in, CUDA tensors out.

One function per entry point of include/lidarreg.h.  PyTorch is used for
device memory and streams only; every computation happens inside
liblidarreg.so.  The reference-facing mirrors (same names and argument
meaning as Experiments/algorithms/*.py) live in
lidarregistration_b200.algorithms and call into this module.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (LrRansacParams, LrRansacStats, SAMPLER_PROSAC, SAMPLER_REPLACE, SAMPLER_UNIFORM,  # noqa: F401
                   SCORE_COUNT, SCORE_MSAC)

DEFAULT_ROUND = 65536


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def to_dev_f32(x):
    """float32 contiguous CUDA tensor from a tensor / ndarray (no copy if already one)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    x = x.detach()
    if not x.is_cuda:
        x = x.to(_dev(), non_blocking=True)
    return x.to(torch.float32).contiguous()


def to_dev_i64(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    x = x.detach()
    if not x.is_cuda:
        x = x.to(_dev(), non_blocking=True)
    return x.to(torch.int64).contiguous()


# ---------------------------------------------------------------- matching
def match_nn(f0, f1, want_2nd=False):
    """lr_match_nn: (idx1[N], idx1_2nd[N] | None), int64 CUDA tensors."""
    f0, f1 = to_dev_f32(f0), to_dev_f32(f1)
    N, D = f0.shape
    M = f1.shape[0]
    assert f1.shape[1] == D
    idx1 = torch.empty(N, dtype=torch.int64, device=f0.device)
    idx2 = torch.empty(N, dtype=torch.int64, device=f0.device) if want_2nd else None
    if N == 0:
        return idx1, idx2
    rc = _lib.lib().lr_match_nn(_lib.ptr(f0), ctypes.c_int64(N), _lib.ptr(f1), ctypes.c_int64(M), int(D),
                                _lib.ptr(idx1), _lib.ptr(idx2), _lib.stream_ptr())
    _lib.check(rc, "lr_match_nn")
    return idx1, idx2


def match_mutual(f0, f1, idx1):
    """lr_match_mutual: mutual pairs (i, j) sorted by i, int64 CUDA tensors."""
    f0, f1, idx1 = to_dev_f32(f0), to_dev_f32(f1), to_dev_i64(idx1)
    N, D = f0.shape
    M = f1.shape[0]
    out_i = torch.empty(N, dtype=torch.int64, device=f0.device)
    out_j = torch.empty(N, dtype=torch.int64, device=f0.device)
    K = torch.zeros(1, dtype=torch.int64, device=f0.device)
    if N == 0:
        return out_i, out_j
    rc = _lib.lib().lr_match_mutual(_lib.ptr(f0), ctypes.c_int64(N), _lib.ptr(f1), ctypes.c_int64(M), int(D),
                                    _lib.ptr(idx1), _lib.ptr(out_i), _lib.ptr(out_j), _lib.ptr(K),
                                    _lib.stream_ptr())
    _lib.check(rc, "lr_match_mutual")
    k = int(K.item())
    return out_i[:k], out_j[:k]


def match_ratio(f0, f1, i0, i1, i2):
    f0, f1 = to_dev_f32(f0), to_dev_f32(f1)
    i0, i1, i2 = to_dev_i64(i0), to_dev_i64(i1), to_dev_i64(i2)
    K = i0.shape[0]
    out = torch.empty(K, dtype=torch.float32, device=f0.device)
    if K == 0:
        return out
    rc = _lib.lib().lr_match_ratio(_lib.ptr(f0), _lib.ptr(f1), int(f0.shape[1]), ctypes.c_int64(K), _lib.ptr(i0),
                                   _lib.ptr(i1), _lib.ptr(i2), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lr_match_ratio")
    return out


def gather_xyz(xyz, idx):
    xyz, idx = to_dev_f32(xyz), to_dev_i64(idx)
    K = idx.shape[0]
    out = torch.empty((K, 3), dtype=torch.float32, device=xyz.device)
    if K == 0:
        return out
    rc = _lib.lib().lr_gather_xyz(_lib.ptr(xyz), _lib.ptr(idx), ctypes.c_int64(K), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lr_gather_xyz")
    return out


def gpf_filter(ratio, is_bb, xyz0, idx0, grid_wid, total_num):
    """lr_gpf_filter -> (keep[n] bool CUDA, norm[n] fp32 CUDA): the Grid-Prioritised Filter's selection on the device."""
    ratio, xyz0 = to_dev_f32(ratio), to_dev_f32(xyz0)
    n = ratio.shape[0]
    keep = torch.zeros(n, dtype=torch.uint8, device=ratio.device)
    norm = torch.empty(n, dtype=torch.float32, device=ratio.device)
    if n == 0:
        return keep.bool(), norm
    if is_bb is not None:
        is_bb = is_bb.to(ratio.device).to(torch.uint8).contiguous()
    idx0 = to_dev_i64(idx0) if idx0 is not None else None
    rc = _lib.lib().lr_gpf_filter(_lib.ptr(ratio), _lib.ptr(is_bb), _lib.ptr(xyz0), _lib.ptr(idx0), ctypes.c_int64(n),
                                  int(grid_wid), ctypes.c_double(float(total_num)), _lib.ptr(keep), _lib.ptr(norm),
                                  _lib.stream_ptr())
    _lib.check(rc, "lr_gpf_filter")
    return keep.bool(), norm


# ------------------------------------------------------------------ RANSAC
def make_params(threshold=0.6, confidence=1.0, max_iters=500000, seed=51, sample_size=3,
                sampler=SAMPLER_UNIFORM, use_elc=True, elc_ratio=0.9, round_size=DEFAULT_ROUND, refit=True,
                scoring=SCORE_COUNT, lo_rounds=0, lo_trials=20, lsq_iters=0):
    """scoring = SCORE_MSAC selects GC-RANSAC semantics: MSAC selection, then lo_rounds rounds of
    local optimisation (lo_trials inner draws each) and lsq_iters passes of iterated least squares."""
    p = LrRansacParams()
    p.threshold, p.confidence, p.elc_ratio = float(threshold), float(confidence), float(elc_ratio)
    p.max_iters, p.seed = int(max_iters), int(seed)
    p.sample_size, p.sampler, p.use_elc = int(sample_size), int(sampler), int(bool(use_elc))
    p.round_size, p.refit, p.reserved = int(round_size), int(bool(refit)), 0
    p.scoring, p.lo_rounds, p.lo_trials, p.lsq_iters = int(scoring), int(lo_rounds), int(lo_trials), int(lsq_iters)
    return p


_PINNED_MASK = None  # grow-only pinned staging for mask_on_host (cudaHostAlloc per call costs ~0.1 ms)


def _pinned_mask(n):
    global _PINNED_MASK
    if _PINNED_MASK is None or _PINNED_MASK.numel() < n:
        _PINNED_MASK = torch.empty(max(n, 1 << 16), dtype=torch.uint8, pin_memory=True)
    return _PINNED_MASK[:n]


def _dev_or_pinned_f32(x):
    """A pinned, contiguous fp32 HOST tensor is passed to the library as it is: under unified addressing the
    pack kernel reads it straight over PCIe / C2C (the host->device transfer of the call, no staging copy, no
    extra launches); everything else goes through to_dev_f32."""
    if isinstance(x, torch.Tensor) and not x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() \
            and x.is_pinned() and x.numel() > 0:
        _lib.require_cuda()
        return x.detach()
    return to_dev_f32(x)


def _dev_or_host_f32(x):
    """lr_ransac_rigid_batch takes HOST arrays as they are (contiguous fp32; pinned ones move at full PCIe rate): the
    library copies them in with the copy engine on the pair's lane, under the other lane's kernels."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    if isinstance(x, torch.Tensor) and not x.is_cuda and x.numel() > 0:
        _lib.require_cuda()
        return x.detach().to(torch.float32).contiguous()
    return to_dev_f32(x)


def ransac_rigid(src, tgt, params, want_mask=False, mask_on_host=False):
    """lr_ransac_rigid -> dict(T, T_refit, mask | None, + LrRansacStats fields).

    mask: CUDA bool tensor, or with mask_on_host a numpy bool array: the kernel then writes the inlier bytes
    straight into pinned, device-mapped host memory (unified addressing) and the call's own synchronisation
    covers them -- no conversion kernel, no second copy, no second synchronisation."""
    return _rigid_call("lr_ransac_rigid", src, tgt, params, want_mask, mask_on_host)


def _rigid_call(fn_name, src, tgt, params, want_mask, mask_on_host):
    src, tgt = _dev_or_pinned_f32(src), _dev_or_pinned_f32(tgt)
    n = src.shape[0]
    T = (ctypes.c_double * 16)()
    Tr = (ctypes.c_double * 16)()
    st = LrRansacStats()
    mask = None
    if want_mask:
        mask = _pinned_mask(n) if mask_on_host else torch.empty(n, dtype=torch.uint8, device=_dev())
    mask_ptr = None if mask is None else ctypes.c_void_p(mask.data_ptr())
    rc = getattr(_lib.lib(), fn_name)(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(tgt.data_ptr()),
                                      ctypes.c_int64(n), ctypes.byref(params), T, Tr,
                                      mask_ptr, ctypes.byref(st), _lib.stream_ptr())
    _lib.check(rc, fn_name)
    if mask is not None:
        mask = mask.numpy().astype(bool) if mask_on_host else mask.bool()
    out = dict(T=_lib.T_from16(T), T_refit=_lib.T_from16(Tr), mask=mask)
    out.update(st.as_dict())
    return out


# ------------------------------------------------ hypothesis sharding with the library-owned communicator
def comm_world():
    """(rank, world) of the communicator connected on the current device; world == 0: none"""
    _lib.require_cuda()
    r, w = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().lr_comm_info(ctypes.byref(r), ctypes.byref(w)), "lr_comm_info")
    return int(r.value), int(w.value)


def comm_init(rank, world, all_gather_bytes):
    """lr_comm_init + lr_comm_connect.  `all_gather_bytes(b: bytes) -> [bytes] * world` (rank order) is the
    caller's transport for the 64-byte IPC handles (torch.distributed all_gather_object, MPI, ...)."""
    _lib.require_cuda()
    h = (ctypes.c_uint8 * 64)()
    _lib.check(_lib.lib().lr_comm_init(int(rank), int(world), h), "lr_comm_init")
    handles = all_gather_bytes(bytes(h))
    assert len(handles) == world and all(len(x) == 64 for x in handles)
    buf = (ctypes.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
    _lib.check(_lib.lib().lr_comm_connect(buf), "lr_comm_connect")


def comm_destroy():
    _lib.check(_lib.lib().lr_comm_destroy(), "lr_comm_destroy")


def ransac_rigid_sharded(src, tgt, params, want_mask=False, mask_on_host=False):
    """lr_ransac_rigid_sharded (a collective over the connected communicator) -> the dict ransac_rigid returns"""
    return _rigid_call("lr_ransac_rigid_sharded", src, tgt, params, want_mask, mask_on_host)


def tc_probe(src, tgt, models, threshold=0.6, want_d=True, want_counts=True):
    """lr_ransac_tc_probe -> (d[H, pad128(n), 3] fp32 | None, E[H] fp64, counts[H] int32 | None), CUDA tensors"""
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    if isinstance(models, np.ndarray):
        models = torch.from_numpy(np.ascontiguousarray(models))
    models = models.to(src.device).to(torch.float64).contiguous().reshape(-1, 12)
    n, H = src.shape[0], models.shape[0]
    n_pad = (n + 127) // 128 * 128
    d = torch.empty((H, n_pad, 3), dtype=torch.float32, device=src.device) if want_d else None
    E = torch.empty(H, dtype=torch.float64, device=src.device)
    counts = torch.empty(H, dtype=torch.int32, device=src.device) if want_counts else None
    rc = _lib.lib().lr_ransac_tc_probe(_lib.ptr(src), _lib.ptr(tgt), ctypes.c_int64(n), _lib.ptr(models), ctypes.c_int64(H),
                                       ctypes.c_double(threshold), _lib.ptr(d), _lib.ptr(E), _lib.ptr(counts),
                                       _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_tc_probe")
    return d, E, counts


def ransac_rigid_batch(pairs, params):
    """lr_ransac_rigid_batch: `pairs` = [(src, tgt), ...] of fp32 [n,3] CUDA tensors, or HOST tensors / numpy arrays
    (copied in by the library, pair i + 1 under the kernels of pair i) -> list of dicts as ransac_rigid returns (no
    masks).  Two pairs in flight at a time, one host synchronisation for the batch."""
    pairs = [(_dev_or_host_f32(a), _dev_or_host_f32(b)) for a, b in pairs]
    k = len(pairs)
    if k == 0:
        return []
    srcs = (ctypes.c_void_p * k)(*[a.data_ptr() for a, _ in pairs])
    tgts = (ctypes.c_void_p * k)(*[b.data_ptr() for _, b in pairs])
    ns = (ctypes.c_int64 * k)(*[a.shape[0] for a, _ in pairs])
    T = (ctypes.c_double * (16 * k))()
    Tr = (ctypes.c_double * (16 * k))()
    st = (LrRansacStats * k)()
    rc = _lib.lib().lr_ransac_rigid_batch(srcs, tgts, ns, ctypes.c_int(k), ctypes.byref(params), T, Tr, st,
                                          _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_rigid_batch")
    out = []
    for i in range(k):
        d = dict(T=_lib.T_from16(T[16 * i:16 * i + 16]), T_refit=_lib.T_from16(Tr[16 * i:16 * i + 16]), mask=None)
        d.update(st[i].as_dict())
        out.append(d)
    return out


def ransac_score_samples(src, tgt, samples, threshold=0.6, use_elc=True, elc_ratio=0.9, want_models=False):
    """Fed-sample parity hook -> (counts[H] int32 CUDA, best, models[H,12] | None)."""
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    if isinstance(samples, np.ndarray):
        samples = torch.from_numpy(np.ascontiguousarray(samples))
    samples = samples.to(src.device).to(torch.int32).contiguous()
    H, m = samples.shape
    counts = torch.empty(H, dtype=torch.int32, device=src.device)
    models = torch.empty((H, 12), dtype=torch.float64, device=src.device) if want_models else None
    best = ctypes.c_int64(-1)
    rc = _lib.lib().lr_ransac_score_samples(_lib.ptr(src), _lib.ptr(tgt), ctypes.c_int64(src.shape[0]),
                                            _lib.ptr(samples), ctypes.c_int64(H), int(m), ctypes.c_double(threshold),
                                            int(bool(use_elc)), ctypes.c_double(elc_ratio), _lib.ptr(counts),
                                            _lib.ptr(models), ctypes.byref(best), _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_score_samples")
    return counts, int(best.value), models


def ransac_score_samples_msac(src, tgt, samples, threshold=0.6, use_elc=True, elc_ratio=0.9):
    """Fed-sample parity hook, MSAC flavour -> (scores[H] int64 CUDA, inliers[H] int32 CUDA, best)."""
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    if isinstance(samples, np.ndarray):
        samples = torch.from_numpy(np.ascontiguousarray(samples))
    samples = samples.to(src.device).to(torch.int32).contiguous()
    H, m = samples.shape
    scores = torch.empty(H, dtype=torch.int64, device=src.device)
    inl = torch.empty(H, dtype=torch.int32, device=src.device)
    best = ctypes.c_int64(-1)
    rc = _lib.lib().lr_ransac_score_samples_msac(_lib.ptr(src), _lib.ptr(tgt), ctypes.c_int64(src.shape[0]),
                                                 _lib.ptr(samples), ctypes.c_int64(H), int(m),
                                                 ctypes.c_double(threshold), int(bool(use_elc)),
                                                 ctypes.c_double(elc_ratio), _lib.ptr(scores), _lib.ptr(inl),
                                                 ctypes.byref(best), _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_score_samples_msac")
    return scores, inl, int(best.value)


def ransac_shard(src, tgt, params, id_lo, id_hi, key):
    """lr_ransac_shard: asynchronously max-merge the packed best of ids [id_lo, id_hi) into key[0]."""
    assert key.is_cuda and key.dtype == torch.int64 and key.numel() == 1
    rc = _lib.lib().lr_ransac_shard(_lib.ptr(src), _lib.ptr(tgt), ctypes.c_int64(src.shape[0]), ctypes.byref(params),
                                    ctypes.c_int64(id_lo), ctypes.c_int64(id_hi), _lib.ptr(key), _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_shard")


def ransac_finalize(src, tgt, params, key, want_mask=False):
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    n = src.shape[0]
    T = (ctypes.c_double * 16)()
    Tr = (ctypes.c_double * 16)()
    st = LrRansacStats()
    mask = torch.empty(n, dtype=torch.uint8, device=src.device) if want_mask else None
    rc = _lib.lib().lr_ransac_finalize(_lib.ptr(src), _lib.ptr(tgt), ctypes.c_int64(n), ctypes.byref(params),
                                       ctypes.c_uint64(int(key) & 0xFFFFFFFFFFFFFFFF), T, Tr, _lib.ptr(mask),
                                       ctypes.byref(st), _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_finalize")
    out = dict(T=_lib.T_from16(T), T_refit=_lib.T_from16(Tr), mask=mask.bool() if want_mask else None)
    out.update(st.as_dict())
    return out


def conf_iters(c, n, m, confidence, max_iters):
    """Host-only stopping rule (needs no GPU)."""
    return int(_lib.lib().lr_ransac_conf_iters(int(c), int(n), int(m), float(confidence), int(max_iters)))


def ransac_sample(params, n, id_lo, H):
    _lib.require_cuda()
    out = torch.empty((H, params.sample_size), dtype=torch.int32, device=_dev())
    rc = _lib.lib().lr_ransac_sample(ctypes.byref(params), ctypes.c_int64(n), ctypes.c_int64(id_lo),
                                     ctypes.c_int64(H), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lr_ransac_sample")
    return out


def refit_indexed(xyz0, xyz1, i0, i1, T, threshold=0.6):
    """lr_refit_indexed (FR.py:99-111) -> (T[4,4], inlier count)."""
    xyz0, xyz1, i0, i1 = to_dev_f32(xyz0), to_dev_f32(xyz1), to_dev_i64(i0), to_dev_i64(i1)
    Tin = (ctypes.c_double * 16)(*np.asarray(T, dtype=np.float64).reshape(-1))
    Tout = (ctypes.c_double * 16)()
    cnt = ctypes.c_int64(0)
    rc = _lib.lib().lr_refit_indexed(_lib.ptr(xyz0), _lib.ptr(xyz1), _lib.ptr(i0), _lib.ptr(i1),
                                     ctypes.c_int64(i0.shape[0]), Tin, ctypes.c_double(threshold), Tout,
                                     ctypes.byref(cnt), _lib.stream_ptr())
    _lib.check(rc, "lr_refit_indexed")
    return _lib.T_from16(Tout), int(cnt.value)


def key_unpack(key):
    """packed (count + 1) << 32 | (0xFFFFFFFF - id)  ->  (count, id); (−1, −1) for the empty key"""
    key = int(key) & 0xFFFFFFFFFFFFFFFF
    if key == 0:
        return -1, -1
    return (key >> 32) - 1, 0xFFFFFFFF - (key & 0xFFFFFFFF)


def key_pack(count, hid):
    return ((int(count) + 1) << 32) | (0xFFFFFFFF - int(hid))


# ------------------------------------------------------------- measurement
PROF_SCORE, PROF_GEN, PROF_NN, PROF_RECOUNT, PROF_PACK, PROF_END, PROF_FIN = 0, 1, 2, 3, 4, 5, 6


def prof_enable(on=True):
    _lib.check(_lib.lib().lr_prof_enable(int(bool(on))), "lr_prof_enable")


def prof_read(kind):
    """-> (device milliseconds, launches) of one kernel class since the last read"""
    ms, cnt = ctypes.c_double(0.0), ctypes.c_int64(0)
    _lib.check(_lib.lib().lr_prof_read(int(kind), ctypes.byref(ms), ctypes.byref(cnt)), "lr_prof_read")
    return float(ms.value), int(cnt.value)


def peak_fp32(mode=0):
    _lib.require_cuda()
    v = ctypes.c_double(0.0)
    _lib.check(_lib.lib().lr_peak_fp32(int(mode), ctypes.byref(v)), "lr_peak_fp32")
    return float(v.value)


def match_set_mode(mode):
    """0 = tensor-core sweep + exact re-rank (D == 32, default); 1 = exact CUDA-core sweep."""
    _lib.check(_lib.lib().lr_match_set_mode(int(mode)), "lr_match_set_mode")


def ransac_set_mode(mode):
    """0 = tensor-core inlier sweep (default); 1 = fp32 sweep, every residual in full; 2 = fp32 sweep with the
    first-component early-out (round 1's default).  Identical results."""
    _lib.check(_lib.lib().lr_ransac_set_mode(int(mode)), "lr_ransac_set_mode")


# --------------------------------------------------------------------- ICP
def transform_pad8(xyz, T):
    """lr_transform_pad8 -> [n, 8] fp32 rows (T * xyz, zero padded) for a 3-D lr_match_nn(D = 8)"""
    xyz = to_dev_f32(xyz)
    n = xyz.shape[0]
    out = torch.empty((n, 8), dtype=torch.float32, device=xyz.device)
    Tin = (ctypes.c_double * 16)(*np.asarray(T, dtype=np.float64).reshape(-1))
    rc = _lib.lib().lr_transform_pad8(_lib.ptr(xyz), ctypes.c_int64(n), Tin, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lr_transform_pad8")
    return out


def icp_step(xyz0, xyz1, i1, T, threshold):
    """lr_icp_step over the pairs (xyz0[k], xyz1[i1[k]]) -> (T_new, inlier count, sum of squared residuals)"""
    xyz0, xyz1, i1 = to_dev_f32(xyz0), to_dev_f32(xyz1), to_dev_i64(i1)
    Tin = (ctypes.c_double * 16)(*np.asarray(T, dtype=np.float64).reshape(-1))
    Tout = (ctypes.c_double * 16)()
    cnt, err2 = ctypes.c_int64(0), ctypes.c_double(0.0)
    rc = _lib.lib().lr_icp_step(_lib.ptr(xyz0), _lib.ptr(xyz1), None, _lib.ptr(i1), ctypes.c_int64(i1.shape[0]), Tin,
                                ctypes.c_double(threshold), Tout, ctypes.byref(cnt), ctypes.byref(err2),
                                _lib.stream_ptr())
    _lib.check(rc, "lr_icp_step")
    return _lib.T_from16(Tout), int(cnt.value), float(err2.value)


def nn3d_radius(src, tgt, T, radius):
    """lr_nn3d_radius -> (idx[n] int64 CUDA, d2[n] fp64 CUDA): nearest row of tgt to T * src[i] inside the radius, -1 = none"""
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    n, m = src.shape[0], tgt.shape[0]
    idx = torch.empty(n, dtype=torch.int64, device=src.device)
    d2 = torch.empty(n, dtype=torch.float64, device=src.device)
    Tin = (ctypes.c_double * 16)(*np.asarray(T, dtype=np.float64).reshape(-1))
    rc = _lib.lib().lr_nn3d_radius(_lib.ptr(src), ctypes.c_int64(n), _lib.ptr(tgt), ctypes.c_int64(m), Tin,
                                   ctypes.c_double(radius), _lib.ptr(idx), _lib.ptr(d2), _lib.stream_ptr())
    _lib.check(rc, "lr_nn3d_radius")
    return idx, d2


def icp_refine(src, tgt, max_dist, T_init=None, max_iteration=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """lr_icp_refine: the whole point-to-point ICP on the device -> (T[4,4], fitness, inlier_rmse, iterations)"""
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    Tin = None if T_init is None else (ctypes.c_double * 16)(*np.asarray(T_init, dtype=np.float64).reshape(-1))
    Tout = (ctypes.c_double * 16)()
    fit, rmse, it = ctypes.c_double(0.0), ctypes.c_double(0.0), ctypes.c_int(0)
    rc = _lib.lib().lr_icp_refine(_lib.ptr(src) if src.shape[0] else None, ctypes.c_int64(src.shape[0]),
                                  _lib.ptr(tgt) if tgt.shape[0] else None, ctypes.c_int64(tgt.shape[0]),
                                  ctypes.c_double(max_dist), Tin, int(max_iteration), ctypes.c_double(rel_fitness),
                                  ctypes.c_double(rel_rmse), Tout, ctypes.byref(fit), ctypes.byref(rmse), ctypes.byref(it),
                                  _lib.stream_ptr())
    _lib.check(rc, "lr_icp_refine")
    return _lib.T_from16(Tout), float(fit.value), float(rmse.value), int(it.value)


# ------------------------------------------------------- PointDSC seed scoring
def kabsch_weighted_batch(A, B, w=None):
    """lr_kabsch_weighted_batch: A, B [S,k,3], w [S,k] | None -> [S,4,4] fp64 CUDA tensor"""
    A, B = to_dev_f32(A), to_dev_f32(B)
    w = None if w is None else to_dev_f32(w)
    S, k = int(A.shape[0]), int(A.shape[1])
    out = torch.empty((S, 4, 4), dtype=torch.float64, device=A.device)
    if S == 0:
        return out
    rc = _lib.lib().lr_kabsch_weighted_batch(_lib.ptr(A), _lib.ptr(B), _lib.ptr(w), ctypes.c_int64(S), k, _lib.ptr(out),
                                             _lib.stream_ptr())
    _lib.check(rc, "lr_kabsch_weighted_batch")
    return out


def seeds_score(src, tgt, models, threshold, want_labels=True, want_refit=False):
    """lr_seeds_score: models [S,4,4] -> dict(counts[S] int32 CUDA, best, best_count, T, labels[n] bool CUDA | None,
    T_refit | None)"""
    src, tgt = to_dev_f32(src), to_dev_f32(tgt)
    if isinstance(models, np.ndarray):
        models = torch.from_numpy(np.ascontiguousarray(models))
    models = models.detach().to(src.device).to(torch.float64).reshape(-1, 16).contiguous()
    n, S = int(src.shape[0]), int(models.shape[0])
    counts = torch.empty(S, dtype=torch.int32, device=src.device)
    labels = torch.empty(n, dtype=torch.uint8, device=src.device) if want_labels else None
    T = (ctypes.c_double * 16)()
    Tr = (ctypes.c_double * 16)() if want_refit else None
    best, bc = ctypes.c_int64(-1), ctypes.c_int64(-1)
    rc = _lib.lib().lr_seeds_score(_lib.ptr(src), _lib.ptr(tgt), ctypes.c_int64(n), _lib.ptr(models), ctypes.c_int64(S),
                                   ctypes.c_double(threshold), _lib.ptr(counts), _lib.ptr(labels), ctypes.byref(best),
                                   ctypes.byref(bc), T, Tr, _lib.stream_ptr())
    _lib.check(rc, "lr_seeds_score")
    return dict(counts=counts, best=int(best.value), best_count=int(bc.value), T=_lib.T_from16(T),
                labels=labels.bool() if want_labels else None, T_refit=_lib.T_from16(Tr) if want_refit else None)
