"""Inputs of the gpf_ref_ieee fixture, regenerated from seeds on both sides (make_golden.py and the GPU test): the
fixture then only stores the reference's outputs.  numpy's default_rng streams are stable across releases."""
import numpy as np

CASES = [dict(N=3000, M=3200, planted=600, phi=0.5, seed=53, bb_first=False, cap=10 ** 9),
         dict(N=6000, M=5000, planted=1200, phi=2.0, seed=55, bb_first=False, cap=10 ** 9),  # the default phi
         dict(N=5000, M=5000, planted=2000, phi=1.0, seed=56, bb_first=False, cap=10 ** 9),
         dict(N=4000, M=4500, planted=1800, phi=2.0, seed=57, bb_first=True, cap=700)]       # TEASER variant


def unit(x):
    x = x.astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def make(cs):
    rng = np.random.default_rng(cs["seed"])
    N, M, D = cs["N"], cs["M"], 32
    f0 = unit(rng.standard_normal((N, D)))
    f1 = unit(rng.standard_normal((M, D)))
    k = cs["planted"]
    f1[:k] = unit(f0[:k] + 0.1 * rng.standard_normal((k, D)).astype(np.float32))
    # clustered source cloud: cells of very different population, some empty
    xyz0 = (rng.standard_normal((N, 3)) * np.array([25.0, 40.0, 2.0])).astype(np.float32)
    xyz0[: N // 5] = rng.uniform(-70, 70, (N // 5, 3)).astype(np.float32)
    return f0, f1, xyz0
