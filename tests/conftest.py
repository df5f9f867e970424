import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def matching_golden(golden_dir):
    import numpy as np
    z = np.load(os.path.join(golden_dir, "matching_ref.npz"))
    cases = {}
    for key in z.files:
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = z[key]
    return cases


@pytest.fixture(scope="session")
def kabsch_golden(golden_dir):
    import numpy as np
    return np.load(os.path.join(golden_dir, "kabsch_ref.npz"))


@pytest.fixture(scope="session")
def refit_golden(golden_dir):
    """24 cases of the reference's weighted_procrustes (DGR/util/procrustes.py:34-56) with 0/1 weights = the inlier
    mask of a coarse model at 0.6 m: the refit step of FR.py:99-111 (tests/golden/make_golden.py::refit_cases)"""
    import numpy as np
    z = np.load(os.path.join(golden_dir, "refit_ref.npz"))
    cases = {}
    for key in z.files:
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = z[key]
    return [cases[k] for k in sorted(cases)]


@pytest.fixture(scope="session")
def seeds_golden(golden_dir):
    """4 runs of the reference's PointDSC.cal_seed_trans (Experiments/models/PointDSC.py:234-336, unmodified, CPU):
    neighbourhoods + weights handed to rigid_transform_3d, per-seed transforms, fitness, best transform, final labels
    (tests/golden/make_golden.py::seeds_cases)"""
    import numpy as np
    z = np.load(os.path.join(golden_dir, "seeds_ref.npz"))
    thr = float(z["threshold"])
    cases = []
    for c in range(int(z["num_cases"])):
        cases.append({k[len("c%d_" % c):]: z[k] for k in z.files if k.startswith("c%d_" % c)})
        cases[-1]["threshold"] = thr
    return cases
