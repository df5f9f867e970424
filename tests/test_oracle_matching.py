"""The oracle's matching half against the reference's own outputs (golden fixtures made by
tests/golden/make_golden.py from /root/reference/Experiments/algorithms/matching.py)."""
import numpy as np
import pytest

from oracle import lr_oracle as O


def test_find_nn_matches_reference(matching_golden):
    for name, c in matching_golden.items():
        _, i1, i2 = O.find_nn(c["f0"], c["f1"], return_2nd=True)
        assert np.array_equal(i1, c["idx1"]), name
        assert np.array_equal(i2, c["idx2"]), name


def test_mutual_matches_reference(matching_golden):
    for name, c in matching_golden.items():
        mi, mj = O.nn_to_mutual(c["f0"], c["f1"], c["idx1"])
        assert np.array_equal(mi, c["mut_i"]) and np.array_equal(mj, c["mut_j"]), name
        assert np.all(np.diff(mi) > 0)  # sorted by idx0 (coalesce order, matching.py:67-87)


def test_ratio_matches_reference(matching_golden):
    # torch's vectorised CPU sqrt is not correctly rounded in ~0.5 % of values, so the fp32 ratio
    # is pinned to 1 ulp rather than bit-exactly (indices above are bit-exact)
    for name, c in matching_golden.items():
        r = O.ratio(c["f0"], c["f1"], c["mut_i"], c["mut_j"], c["mut_2nd"])
        ref = c["ratio"]
        ok = np.isfinite(ref)
        assert np.allclose(r[ok], ref[ok], rtol=3e-7, atol=0), name


def test_ties_pick_lowest_index():
    rng = np.random.default_rng(3)
    f1 = rng.standard_normal((50, 32)).astype(np.float32)
    f1[40] = f1[7]
    f1[20] = f1[7]
    f0 = f1[[7]].copy()
    _, i1, i2 = O.find_nn(f0, f1, return_2nd=True)
    assert i1[0] == 7 and i2[0] == 20


def test_single_target_second_is_zero():
    f0 = np.ones((3, 32), np.float32)
    f1 = np.ones((1, 32), np.float32)
    _, i1, i2 = O.find_nn(f0, f1, return_2nd=True)
    assert np.all(i1 == 0) and np.all(i2 == 0)  # all-inf row after masking -> index 0


@pytest.mark.parametrize("n,m", [(300, 400), (513, 129)])
def test_mutual_is_symmetric_property(n, m):
    rng = np.random.default_rng(n)
    f0 = rng.standard_normal((n, 32)).astype(np.float32)
    f1 = rng.standard_normal((m, 32)).astype(np.float32)
    _, a, _ = O.find_nn(f0, f1)
    _, b, _ = O.find_nn(f1, f0)
    mi, mj = O.nn_to_mutual(f0, f1, a)
    ni, nj = O.nn_to_mutual(f1, f0, b)
    assert set(zip(mi.tolist(), mj.tolist())) == set(zip(nj.tolist(), ni.tolist()))
