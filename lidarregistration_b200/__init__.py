"""lidarregistration_b200 -- B200-native robust-registration hot path.

FCGF-feature correspondence search (NN / mutual NN) followed by RANSAC rigid
motion estimation, as hand-written sm_100a CUDA behind the C ABI of
include/lidarreg.h.  `algorithms` mirrors the reference's
Experiments/algorithms interface (FR, find_nn, nn_to_mutual, GC_RANSAC, ...);
`engine` is the device-resident front-end; `parallel` shards hypotheses or
pairs over GPUs.  There is no CPU implementation in this package.
"""
__version__ = "0.1.0"

from . import _lib, build, engine  # noqa: F401

