"""The C-ABI shared library loads without a GPU and exports every symbol include/lidarreg.h declares."""
import os
import re

import pytest

from lidarregistration_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lidarreg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in lidarreg.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms
    assert L.lr_version() >= 100


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(_lib.LrRansacParams) == 3 * 8 + 2 * 8 + 10 * 4
    assert ctypes.sizeof(_lib.LrRansacStats) == 9 * 8 + 2 * 4


def test_host_only_entry_points():
    from lidarregistration_b200 import engine
    from oracle import lr_oracle as O
    for c, n, m, conf in [(9000, 30000, 3, 0.9995), (1, 30000, 3, 0.999), (300, 30000, 4, 0.9995), (0, 10, 3, 0.5)]:
        assert engine.conf_iters(c, n, m, conf, 10**6) == O.conf_iters(c, n, m, conf, 10**6)
    assert engine.key_unpack(engine.key_pack(1234, 77)) == (1234, 77)
    assert engine.key_unpack(0) == (-1, -1)
    # ties: lower id wins under MAX
    assert engine.key_pack(10, 3) > engine.key_pack(10, 4) > engine.key_pack(9, 0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lidarregistration_b200 import engine
    with pytest.raises(RuntimeError):
        engine.match_nn(torch.zeros(4, 32), torch.zeros(4, 32))
    with pytest.raises(RuntimeError):
        engine.ransac_rigid(torch.zeros(4, 3), torch.zeros(4, 3), engine.make_params())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lidarregistration_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "lr_oracle" not in txt.replace("oracle/lr_oracle.c", "") and "import oracle" not in txt \
                    and "from oracle" not in txt, f


def test_hypothesis_sharding_refuses_msac_scoring():
    """host-side guard (no GPU needed): the packed (count, id) key cannot carry an MSAC score"""
    import numpy as np
    import pytest
    from lidarregistration_b200 import engine, parallel
    p = engine.make_params(scoring=engine.SCORE_MSAC, lo_rounds=2)
    with pytest.raises(ValueError, match="count scoring only"):
        parallel.ransac_rigid_sharded(np.zeros((10, 3), np.float32), np.zeros((10, 3), np.float32), p)


def test_gc_options_mapping():
    """--GC_scoring / --GC_LO -> native parameters (GC_RANSAC.py:12-37, gcransac_python.cpp:418-423, 511-521)"""
    import sys
    import pytest
    import lidarregistration_b200.algorithms  # noqa: F401
    G = sys.modules["lidarregistration_b200.algorithms.GC_RANSAC"]
    with pytest.warns(UserWarning, match="GC_LO False has no effect"):
        assert G.gc_options(None, True) == dict(scoring=0) == G.gc_options("count", False)
    assert G.gc_options("MSAC", True) == dict(scoring=1, lo_rounds=10, lo_trials=20, lsq_iters=10)
    # --GC_LO False only switches the graph-cut rounds off; the finishing least squares still runs (:518-521, App. A)
    assert G.gc_options("msac", False) == dict(scoring=1, lo_rounds=0, lo_trials=20, lsq_iters=10)
    # --fast_rejection NONE: the no-preemption branch ignores --GC_LO and allows 50 inner draws (:570-592)
    assert G.gc_options("MSAC", False, preemption=False) == dict(scoring=1, lo_rounds=10, lo_trials=50, lsq_iters=10)
    with pytest.raises(ValueError):
        G.gc_options("LMEDS", True)
    with pytest.raises(NotImplementedError):
        G.gc_options("MSAC", True, spatial_coherence_weight=0.1)


def test_metrics_match_reference_transformation_loss(golden_dir):
    """RE / TE / success (stats columns 0-2) against the reference's TransformationLoss.forward
    (Experiments/libs/loss.py:44-51, run by tests/golden/make_golden.py::metrics_cases; fp32 on its side)."""
    import os
    import numpy as np
    from lidarregistration_b200 import metrics
    g = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    for T, Tg, re, te, ok in zip(g["T"], g["T_gt"], g["RE"], g["TE"], g["ok"]):
        assert abs(metrics.rotation_error_deg(T, Tg) - re) < 2e-2 + 1e-4 * re   # acos near 1 in fp32 on the reference side
        assert abs(metrics.translation_error_cm(T, Tg) - te) < 1e-2 + 1e-5 * te
        near = abs(re - 5.0) < 0.05 or abs(te - 60.0) < 0.05
        assert near or metrics.registration_success(T, Tg) == bool(ok)


def test_mode_switches_validate_their_argument():
    """host-only entries of the ABI: bad arguments come back as LR_ERR_ARG with a message, nothing throws"""
    from lidarregistration_b200 import _lib
    L = _lib.lib()
    assert L.lr_ransac_set_mode(0) == 0 and L.lr_ransac_set_mode(1) == 0 and L.lr_ransac_set_mode(0) == 0
    assert L.lr_ransac_set_mode(7) == 1 and b"mode" in L.lr_last_error()
    assert L.lr_match_set_mode(0) == 0
    assert L.lr_match_set_mode(99) != 0 and len(L.lr_last_error()) > 0
    assert L.lr_match_set_mode(0) == 0


def test_f4_entries_validate_their_arguments_before_touching_the_device():
    """lr_icp_refine / lr_nn3d_radius / lr_seeds_score / lr_kabsch_weighted_batch and the measurement switches: bad
    arguments come back as LR_ERR_ARG with a message (checked before any CUDA call, so this runs without a GPU)"""
    import ctypes
    from lidarregistration_b200 import _lib
    L = _lib.lib()
    T = (ctypes.c_double * 16)()
    i64, dbl = ctypes.c_int64, ctypes.c_double
    buf = ctypes.c_void_p(0x1000)  # never dereferenced: the calls below fail their argument checks first
    assert L.lr_icp_refine(buf, i64(10), buf, i64(10), dbl(0.0), None, 30, dbl(1e-6), dbl(1e-6), T, None, None, None, None) == 1
    assert b"max_dist" in L.lr_last_error()
    assert L.lr_icp_refine(buf, i64(10), buf, i64(10), dbl(0.6), None, -1, dbl(1e-6), dbl(1e-6), T, None, None, None, None) == 1
    assert L.lr_icp_refine(buf, i64(10), buf, i64(10), dbl(0.6), None, 30, dbl(1e-6), dbl(1e-6), None, None, None, None, None) == 1
    assert b"T_out" in L.lr_last_error()
    assert L.lr_nn3d_radius(buf, i64(10), buf, i64(10), None, dbl(-1.0), buf, None, None) == 1 and b"radius" in L.lr_last_error()
    assert L.lr_nn3d_radius(buf, i64(-3), buf, i64(10), None, dbl(0.6), buf, None, None) == 1
    assert L.lr_seeds_score(buf, buf, i64(0), buf, i64(5), dbl(0.6), None, None, None, None, T, None, None) == 1
    assert L.lr_seeds_score(buf, buf, i64(100), buf, i64((1 << 20) + 1), dbl(0.6), None, None, None, None, T, None, None) == 1
    assert L.lr_seeds_score(buf, buf, i64(100), buf, i64(5), dbl(0.0), None, None, None, None, T, None, None) == 1
    assert L.lr_seeds_score(buf, buf, i64(100), None, i64(5), dbl(0.6), None, None, None, None, T, None, None) == 1
    assert L.lr_kabsch_weighted_batch(buf, buf, None, i64(-1), 4, buf, None) == 1
    assert L.lr_kabsch_weighted_batch(buf, buf, None, i64(0), 4, None, None) == 0   # nothing to do
    assert L.lr_debug_slice(0, 1) == 0 and L.lr_debug_slice(3, 2) == 1 and L.lr_debug_slice(0, 0) == 1
    assert L.lr_debug_pdl(0) == 0 and L.lr_debug_pdl(1) == 0
