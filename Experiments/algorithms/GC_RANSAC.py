"""Drop-in location of the reference's Experiments/algorithms/GC_RANSAC.py: re-exports the B200
implementation, so `from algorithms.GC_RANSAC import ...` in Experiments/test.py keeps working."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidarregistration_b200.algorithms.GC_RANSAC import *  # noqa: F401,F403,E402
