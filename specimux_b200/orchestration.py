"""Orchestration around the GPU matching path (mirror of the reference's orchestration.py seams).

setup_match_parameters (reference orchestration.py:548-641) keeps its semantics; its O(B^2)
edlib NW loop runs as one batched GPU kernel (smx_pairwise_nw) and no Bloom filter is ever built
(the prefilter's observable behaviour is emulated by the barcode kernel).
"""
import ctypes as C
import itertools
import logging
import math
import os
import timeit
from datetime import datetime
from typing import Dict, List, Tuple

import numpy as np

from . import _lib
from .constants import Primer
from .databases import BloomEmulationPrefilter, PassthroughPrefilter
from .demultiplex import process_sequences
from .io_utils import (OutputManager, cleanup_empty_directories, open_sequence_file, output_write_operation,
                       read_primers_file, read_specimen_file)
from .models import MatchParameters
from .seqio import reverse_complement


def pairwise_nw(seqs: List[str], device: int = 0) -> np.ndarray:
    """All-pairs global edit distance on the GPU (replaces orchestration.py:549-555)."""
    lib = _lib.load()
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    out = np.zeros((n, n), dtype=np.int32)
    _lib.check(lib.smx_pairwise_nw(device, "".join(seqs).encode("ascii"), _lib.ptr(off, _lib.u32p), n,
                                   _lib.ptr(out, _lib.i32p)))
    return out


def _min_pairwise(seqs: List[str], device: int):
    if len(seqs) <= 1:
        return None
    d = pairwise_nw(seqs, device)
    iu = np.triu_indices(len(seqs), k=1)
    return int(d[iu].min())


def _bp_adjusted_length(primer: str) -> float:
    """reference: orchestration.py:564-570."""
    score = 0
    for b in primer:
        if b in "ACGT":
            score += 3
        elif b in "KMRSWY":
            score += 2
        elif b in "BDHV":
            score += 1
    return score / 3.0


def thresholds_for(specimens, index_edit_distance: int = -1, primer_edit_distance: int = -1,
                   device: int = 0, diagnostics=False) -> Tuple[int, Dict[str, int]]:
    """(k_idx, {primer sequence: k}) -- reference orchestration.py:572-614."""
    b1s, b2s = {}, {}
    for p in specimens.get_primers(Primer.FWD):
        b1s.update(p.barcodes)
    for p in specimens.get_primers(Primer.REV):
        b2s.update(p.barcodes)
    combined = list(b1s) + [reverse_complement(b) for b in b2s]
    if index_edit_distance != -1:
        k_idx = index_edit_distance
        if diagnostics:
            for desc, seqs in (("Forward Barcodes", list(b1s)), ("Reverse Barcodes", list(b2s)),
                               ("Forward Barcodes + Reverse Complement of Reverse Barcodes", combined)):
                m = _min_pairwise(seqs, device)
                if m is not None:
                    logging.info(f"Minimum edit distance is {m} for {desc}")
    else:
        min_bc = _min_pairwise(combined, device)
        if diagnostics and min_bc is not None:
            logging.info(f"Minimum edit distance is {min_bc} for Forward Barcodes + Reverse Complement of Reverse Barcodes")
        k_idx = math.ceil(min_bc / 2.0)      # raises TypeError on a single barcode, like the reference
    thr = {}
    for p in specimens.get_primers(Primer.FWD) + specimens.get_primers(Primer.REV):
        thr[p.primer] = primer_edit_distance if primer_edit_distance != -1 else int(_bp_adjusted_length(p.primer) / 3)
    return k_idx, thr


def setup_match_parameters(args, specimens, device: int = 0) -> MatchParameters:
    """reference: orchestration.py:548-641."""
    k_idx, thr = thresholds_for(specimens, args.index_edit_distance, args.primer_edit_distance, device,
                                getattr(args, "diagnostics", None))
    logging.info(f"Using Edit Distance Thresholds {k_idx} for barcode indexes")
    for p, pt in thr.items():
        logging.info(f"Using Edit Distance Threshold {pt} for primer {p}")
    logging.info(f"Using dereplication strategy: {args.dereplicate}")
    if args.disable_preorient:
        logging.info("Sequence pre-orientation disabled, may run slower")
    if not args.disable_prefilter:
        if specimens.b_length() > 13:
            logging.warning("Barcode prefilter not tested for barcodes longer than 13 nt.  You may need to use --disable-prefilter")
        if k_idx > 3:
            logging.warning("Barcode prefilter not tested for edit distance greater than 3.  You may need to use --disable-prefilter")
        logging.info("Using Bloom Filter optimization for barcode matching")
    else:
        logging.info("Barcode prefiltering disabled, may run slower")
    return MatchParameters(thr, k_idx, args.search_len, not args.disable_preorient)
