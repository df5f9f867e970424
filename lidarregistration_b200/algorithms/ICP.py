"""Point-to-point ICP refinement: the step right after FR() in the reference,
`o3d.pipelines.registration.registration_icp(src, tgt, 0.6, T_init, TransformationEstimationPointToPoint())`
(Experiments/test.py:183-188; Open3D defaults: 30 iterations, relative fitness / rmse 1e-6).

The whole refinement runs on the device (lr_icp_refine): the target is binned once into a hashed uniform grid of
max_correspondence_distance-sized cells -- the role of Open3D's KD-tree --, every iteration is one kernel (transform,
nearest target inside the distance, sums, Kabsch, stopping rule), all iterations are enqueued up front and there is
one synchronisation per call.  `registration_icp_bruteforce` keeps round 2's first composition (exact fp32 sweep,
lr_match_nn with D = 8, one lr_icp_step per iteration) for A/B timing.
"""
import numpy as np

from .. import engine


class RegistrationResult:
    """the fields of o3d.pipelines.registration.RegistrationResult the caller reads"""

    def __init__(self, transformation, fitness, inlier_rmse, iterations):
        self.transformation, self.fitness, self.inlier_rmse, self.iterations = transformation, fitness, inlier_rmse, iterations


def _cloud(x):
    return engine.to_dev_f32(np.asarray(x.points, dtype=np.float32) if hasattr(x, "points") else x)


def registration_icp(source, target, max_correspondence_distance, init=None, max_iteration=30, relative_fitness=1e-6,
                     relative_rmse=1e-6):
    """source / target: [n,3] / [m,3] arrays, tensors or point clouds with `.points`."""
    T, fitness, rmse, it = engine.icp_refine(_cloud(source), _cloud(target), max_correspondence_distance, init, max_iteration,
                                             relative_fitness, relative_rmse)
    return RegistrationResult(T, fitness, rmse, it)


def registration_icp_bruteforce(source, target, max_correspondence_distance, init=None, max_iteration=30,
                                relative_fitness=1e-6, relative_rmse=1e-6):
    """the same loop with the nearest neighbour taken from the exact fp32 sweep over ALL target points"""
    src, tgt = _cloud(source), _cloud(target)
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
