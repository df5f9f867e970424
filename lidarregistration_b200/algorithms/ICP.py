"""Point-to-point ICP refinement: the step right after FR() in the reference,
`o3d.pipelines.registration.registration_icp(src, tgt, 0.6, T_init, TransformationEstimationPointToPoint())`
(Experiments/test.py:183-188; Open3D defaults: 30 iterations, relative fitness / rmse 1e-6).

Composed from the hot path's own kernels: the source is transformed and padded to 8 floats
(lr_transform_pad8), its nearest target point comes from the exact fp32 sweep (lr_match_nn, D = 8),
and one lr_icp_step keeps the pairs closer than the threshold, accumulates fitness / rmse and
solves Kabsch on them.
"""
import numpy as np

from .. import engine


class RegistrationResult:
    """the fields of o3d.pipelines.registration.RegistrationResult the caller reads"""

    def __init__(self, transformation, fitness, inlier_rmse, iterations):
        self.transformation, self.fitness, self.inlier_rmse, self.iterations = transformation, fitness, inlier_rmse, iterations


def registration_icp(source, target, max_correspondence_distance, init=None, max_iteration=30, relative_fitness=1e-6,
                     relative_rmse=1e-6):
    """source / target: [n,3] / [m,3] arrays, tensors or point clouds with `.points`."""
    src = engine.to_dev_f32(np.asarray(source.points, dtype=np.float32) if hasattr(source, "points") else source)
    tgt = engine.to_dev_f32(np.asarray(target.points, dtype=np.float32) if hasattr(target, "points") else target)
    T = np.eye(4) if init is None else np.asarray(init, dtype=np.float64).copy()
    n = src.shape[0]
    tgt8 = engine.transform_pad8(tgt, np.eye(4))

    def evaluate(T):
        idx, _ = engine.match_nn(engine.transform_pad8(src, T), tgt8)
        T_new, cnt, err2 = engine.icp_step(src, tgt, idx, T, max_correspondence_distance)
        fitness = cnt / n if n else 0.0
        rmse = float(np.sqrt(err2 / cnt)) if cnt else 0.0
        return T_new, fitness, rmse

    T_next, fitness, rmse = evaluate(T)
    it = 0
    for it in range(1, max_iteration + 1):
        T = T_next  # Kabsch over the current correspondences (absolute transform of the original source)
        T_next, f2, r2 = evaluate(T)
        done = abs(fitness - f2) < relative_fitness and abs(rmse - r2) < relative_rmse
        fitness, rmse = f2, r2
        if done:
            break
    return RegistrationResult(T, fitness, rmse, it)
