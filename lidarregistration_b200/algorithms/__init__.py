"""Mirror of the reference's Experiments/algorithms package for the RANSAC path."""
from .FR import FR, RANSAC_registration, PointCloud  # noqa: F401
from .GC_RANSAC import GC_RANSAC, findRigidTransform  # noqa: F401
from .matching import (Grid_Prioritized_Filter, calc_distance_ratio_in_feature_space, find_2nn, find_nn,  # noqa: F401
                       mark_best_buddies, measure_inlier_ratio, nn_to_mutual, torch_intersect)
from .ICP import registration_icp, registration_icp_bruteforce  # noqa: F401
from .seeds import score_seeds, seedwise_transforms  # noqa: F401
