"""Drop-in location of the reference's Experiments/algorithms/matching.py: re-exports the B200
implementation, so `from algorithms.matching import ...` in Experiments/test.py keeps working."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidarregistration_b200.algorithms.matching import *  # noqa: F401,F403,E402
