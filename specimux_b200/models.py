"""Data contract of the matching path (mirror of the reference's src/specimux/models.py).

Only the types that cross the process_sequences boundary exist here; the per-candidate state
(CandidateMatch / AlignmentResult in the reference, models.py:34-328) lives on the GPU.
"""
from typing import Dict, List, NamedTuple, Optional, Tuple

from .constants import Primer, ResolutionType
from .seqio import reverse_complement


class PrimerInfo:
    """reference: models.py:20-31.  `barcodes` keeps first-appearance order (the reference uses a
    `set`, whose iteration order depends on PYTHONHASHSEED -- SURVEY.md Q2)."""

    def __init__(self, name: str, seq: str, direction: Primer, pools: List[str], file_index: int = 0):
        self.name = name
        self.primer = seq.upper()
        self.direction = direction
        self.primer_rc = reverse_complement(self.primer)
        self.barcodes: Dict[str, None] = {}
        self.specimens = set()
        self.pools = pools
        self.file_index = file_index

    def __repr__(self):
        return "PrimerInfo(%r, %r)" % (self.name, self.primer)


class MatchParameters:
    """reference: models.py:331-338."""

    def __init__(self, max_dist_primers: Dict[str, int], max_dist_index: int, search_len: int, preorient: bool):
        self.max_dist_primers = max_dist_primers
        self.max_dist_index = max_dist_index
        self.search_len = search_len
        self.preorient = preorient


class WriteOperation(NamedTuple):
    """reference: models.py:341-357."""
    sample_id: str
    seq_id: str
    distance_code: str
    sequence: str
    quality_sequence: str
    quality_scores: List[int]
    p1_location: Optional[Tuple[int, int]]
    p2_location: Optional[Tuple[int, int]]
    b1_location: Optional[Tuple[int, int]]
    b2_location: Optional[Tuple[int, int]]
    primer_pool: str
    p1_name: str
    p2_name: str
    resolution_type: ResolutionType
    trace_sequence_id: Optional[str] = None


class SequenceBatch(NamedTuple):
    """reference: models.py:360-365."""
    seq_number: int
    seq_records: List
    parameters: MatchParameters
    start_idx: int


class WorkerException(Exception):
    pass
