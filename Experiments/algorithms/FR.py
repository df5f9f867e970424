"""Drop-in location of the reference's Experiments/algorithms/FR.py: re-exports the B200
implementation, so `from algorithms.FR import ...` in Experiments/test.py keeps working."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lidarregistration_b200.algorithms.FR import *  # noqa: F401,F403,E402
