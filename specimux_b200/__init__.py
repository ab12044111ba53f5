"""specimux_b200 -- B200-native drop-in for specimux's read-matching path.

The matching (primer HW search in both orientations, barcode SHW search in the flank,
best-match selection and dereplication) runs as hand-written CUDA for sm_100a behind the
C ABI in include/specimux_b200.h; this package is the batching layer that mirrors the
reference's Python interface (src/specimux/__init__.py:14-25 of the reference).
There is no CPU fallback: without the built library or without a GPU the matching raises.
"""
__version__ = "0.7.0+b200.1"

from .constants import (AlignMode, Barcode, MultipleMatchStrategy, Orientation, Primer, ResolutionType,  # noqa: F401
                        SampleId, TrimMode)
from .models import MatchParameters, PrimerInfo, SequenceBatch, WorkerException, WriteOperation  # noqa: F401
from .databases import PrimerDatabase, Specimens  # noqa: F401

__all__ = ["__version__", "PrimerDatabase", "Specimens", "MatchParameters", "PrimerInfo", "WriteOperation",
           "TrimMode", "MultipleMatchStrategy", "ResolutionType", "Primer", "Barcode", "Orientation",
           "SampleId", "AlignMode"]
