"""Stand-in for the `edlib` Python module (TEST INFRASTRUCTURE ONLY)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import aligner as _al  # noqa: E402

CALLS = {"HW": 0, "SHW": 0, "NW": 0}


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
    CALLS[mode] = CALLS.get(mode, 0) + 1
    return _al.align(str(query), str(target), mode, task, k, additionalEqualities)
