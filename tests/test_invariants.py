"""Numerical invariants the CUDA kernels rely on, checked on the CPU with numpy's IEEE arithmetic (no GPU needed)."""
import numpy as np


def fma32(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 values is exact in fp64 (48 bits); the sum is rounded once
    to fp64 and once to fp32, which can differ from a true fma only at fp32 half-way points -- irrelevant to the
    inequalities below, which hold for any monotone rounding"""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def test_early_out_bound_of_the_inlier_sweep():
    """k_score skips a point when |d0| >= c for the whole warp, c = sqrt_ru(hi) * 1.000001f (k_kabsch).  That is
    exact iff every such residual evaluates to rr >= hi, i.e. is counted neither below `lo` nor below `hi`."""
    rng = np.random.default_rng(0)
    for hi in np.float32(rng.uniform(0.2, 0.5, 2000)):
        s = np.sqrt(hi, dtype=np.float32)
        if np.float64(s) * np.float64(s) < np.float64(hi):
            s = np.nextafter(s, np.float32(np.inf))  # round-up square root, as __fsqrt_ru
        c = np.float32(s * np.float32(1.000001))
        # the smallest magnitudes that take the early-out, and a few larger ones
        for d0 in (c, np.nextafter(c, np.float32(np.inf)), np.float32(c * np.float32(1.5)), np.float32(-c)):
            sq = np.float32(d0 * d0)
            assert sq >= hi
            d1, d2 = np.float32(rng.normal(0, 3)), np.float32(rng.normal(0, 3))
            rr = fma32(d2, d2, fma32(d1, d1, sq))  # the sweep's operation order
            assert rr >= sq >= hi
        # and the largest magnitude that does NOT take it may be on either side: it is evaluated in full
        below = np.nextafter(c, np.float32(0))
        assert abs(below) < c


def test_msac_term_quantisation_is_order_free():
    """Integer sums of trunc((1 - r2/tau2) * 65536) do not depend on the order or grouping of the terms, which is
    what lets k_score_msac split a hypothesis's correspondences over CTAs and merge with integer atomics."""
    rng = np.random.default_rng(1)
    tau2 = (1.5 * 0.6) * (1.5 * 0.6)
    r2 = rng.uniform(0, 1.2, 50000)
    terms = np.where(r2 < tau2, ((1.0 - r2 / tau2) * 65536.0).astype(np.int64), 0)
    total = int(terms.sum())
    assert 0 <= terms.min() and terms.max() <= 65536
    for _ in range(5):
        perm = rng.permutation(len(terms))
        parts = np.array_split(terms[perm], int(rng.integers(2, 600)))
        assert sum(int(p.sum()) for p in parts) == total
    # the real-valued score it quantises, summed in two different orders, does differ in the last bits
    w = np.where(r2 < tau2, 1.0 - r2 / tau2, 0.0)
    assert abs(total / 65536.0 - w.sum()) <= (r2 < tau2).sum() / 65536.0


def test_packed_key_orders_by_count_then_lowest_id():
    """(count + 1) << 32 | (0xFFFFFFFF - id): MAX over keys = highest count, ties -> lowest id; 0 = nothing scored"""
    from lidarregistration_b200 import engine
    keys = {(c, i): engine.key_pack(c, i) for c in (0, 1, 7, 30000) for i in (0, 5, 999999, 0xFFFFFFFE)}
    best = max(keys.values())
    assert engine.key_unpack(best) == (30000, 0)
    assert engine.key_pack(7, 5) > engine.key_pack(7, 6) > engine.key_pack(6, 0) > 0
    assert engine.key_unpack(0) == (-1, -1)
    for (c, i), k in keys.items():
        assert engine.key_unpack(k) == (c, i)
