"""ctypes binding of liblidarreg.so (include/lidarreg.h).

This is the thin layer the north star asks for: Python/PyTorch host code on
top, hand-written sm_100a kernels below, a C ABI in between.  There is no
fallback of any kind: if the shared library is missing or a call fails, a
RuntimeError carrying lr_last_error() is raised.
"""
import ctypes
import os

import numpy as np
import torch

from . import build as _build

_LIB = None

c_void_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double


class LrRansacParams(ctypes.Structure):
    _fields_ = [
        ("threshold", ctypes.c_double), ("confidence", ctypes.c_double), ("elc_ratio", ctypes.c_double),
        ("max_iters", ctypes.c_int64), ("seed", ctypes.c_uint64), ("sample_size", ctypes.c_int32),
        ("sampler", ctypes.c_int32), ("use_elc", ctypes.c_int32), ("round_size", ctypes.c_int32),
        ("refit", ctypes.c_int32), ("scoring", ctypes.c_int32), ("lo_rounds", ctypes.c_int32),
        ("lo_trials", ctypes.c_int32), ("lsq_iters", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class LrRansacStats(ctypes.Structure):
    _fields_ = [
        ("iters_run", ctypes.c_int64), ("n_scored", ctypes.c_int64), ("n_rechecked", ctypes.c_int64),
        ("best_id", ctypes.c_int64), ("best_count", ctypes.c_int64), ("refit_count", ctypes.c_int64),
        ("best_score", ctypes.c_int64), ("lo_score", ctypes.c_int64), ("final_score", ctypes.c_int64),
        ("lo_improved", ctypes.c_int32), ("lsq_improved", ctypes.c_int32),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


SAMPLER_UNIFORM, SAMPLER_PROSAC, SAMPLER_REPLACE = 0, 1, 2
SCORE_COUNT, SCORE_MSAC = 0, 1

# every symbol include/lidarreg.h declares (tests check the .so exports them all)
SYMBOLS = [
    "lr_last_error", "lr_version", "lr_device_info", "lr_shutdown", "lr_match_nn", "lr_match_mutual",
    "lr_match_ratio", "lr_gather_xyz", "lr_ransac_rigid", "lr_ransac_rigid_batch", "lr_ransac_score_samples",
    "lr_ransac_score_samples_msac", "lr_ransac_shard",
    "lr_ransac_finalize", "lr_ransac_conf_iters", "lr_ransac_sample", "lr_refit_indexed",
    "lr_prof_enable", "lr_prof_read", "lr_peak_fp32", "lr_match_set_mode", "lr_transform_pad8", "lr_icp_step", "lr_ransac_set_mode",
    "lr_comm_init", "lr_comm_connect", "lr_comm_info", "lr_comm_destroy", "lr_ransac_rigid_sharded", "lr_ransac_tc_probe",
    "lr_gpf_filter", "lr_icp_refine", "lr_nn3d_radius", "lr_kabsch_weighted_batch", "lr_seeds_score",
    "lr_debug_slice", "lr_debug_pdl",
]


def so_path():
    return _build.SO


def lib():
    """Load liblidarreg.so; loud failure if it was never built."""
    global _LIB
    if _LIB is None:
        path = so_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m lidarregistration_b200.build` "
                "(the CUDA extension is the only implementation; there is no CPU fallback)")
        L = ctypes.CDLL(path)
        L.lr_last_error.restype = ctypes.c_char_p
        L.lr_ransac_conf_iters.restype = ctypes.c_int64
        L.lr_ransac_conf_iters.argtypes = [c_i64, c_i64, c_int, c_dbl, c_i64]
        _LIB = L
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().lr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("lidarregistration_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def T_from16(buf):
    return np.array(buf, dtype=np.float64).reshape(4, 4)
