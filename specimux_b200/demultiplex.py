"""process_sequences -- the operator seam of the reference (demultiplex.py:108-212), GPU-backed.

Same signature, same return value: (write_ops in read order then candidate order, total, matched).
Everything between the seam and the WriteOperations -- reverse complement, orientation test,
candidate enumeration, primer HW search, barcode SHW search, scoring, dereplication, specimen
resolution, trim extents (reference demultiplex.py:216-820) -- runs on the GPU through the C ABI;
this function only packs the reads and formats the returned records.
"""
import argparse
from typing import List, Optional, Tuple

from . import _lib
from .constants import ResolutionType, SampleId
from .databases import PassthroughPrefilter
from .engine import Matcher, PackedBatch
from .models import MatchParameters, WriteOperation
from .seqio import get_bases_and_quality, reverse_complement
from .tables import MatchTables

_RES = {r.value: r for r in ResolutionType}
_matcher_cache = {}


def _flags(args, prefilter):
    use_prefilter = prefilter is not None and not isinstance(prefilter, PassthroughPrefilter)
    return (getattr(args, "trim", "barcodes"), getattr(args, "dereplicate", "best"), use_prefilter,
            getattr(args, "min_length", -1), getattr(args, "max_length", -1))


def get_matcher(parameters: MatchParameters, specimens, args, prefilter, device: int = 0, binding=None) -> Matcher:
    """One cached device context per (tables, flags, device)."""
    flags = _flags(args, prefilter)
    key = (id(specimens), id(parameters), flags, device, id(binding))
    m = _matcher_cache.get(key)
    if m is None:
        tables = MatchTables(specimens, parameters, trim=flags[0], dereplicate=flags[1], prefilter=flags[2],
                             min_length=flags[3], max_length=flags[4])
        m = Matcher(tables, device=device, binding=binding)
        _matcher_cache[key] = m
    return m


def _loc(pair):
    a, b = int(pair[0]), int(pair[1])
    return None if a == _lib.NONE else (a, b)


def records_to_write_ops(tables: MatchTables, result, seq_ids, bases, quals, trace_ids=None) -> List[WriteOperation]:
    """smx_record -> WriteOperation (reference: demultiplex.py:30-103 create_write_operation)."""
    ops = []
    rc_cache = {}
    for rec in result.records:
        r = int(rec["read"])
        seq, qual = bases[r], quals[r]
        if rec["reverse"]:
            if r not in rc_cache:
                rc_cache[r] = (reverse_complement(seq), None if qual is None else qual[::-1])
            seq, qual = rc_cache[r]
        if qual is None:
            qual_full = "I" * len(seq)          # get_quality_seq: [40] * len (alignment.py:52-56)
        else:
            qual_full = qual
        s, e = int(rec["trim_start"]), int(rec["trim_end"])
        res = _RES[int(rec["resolution"])]
        d = rec["dist"]
        code = ",".join(str(int(x)) if x >= 0 else "X" for x in d)
        if rec["trim_empty"]:
            sample, pool, p1n, p2n = SampleId.UNKNOWN, "unknown", "unknown", "unknown"
            out_seq, out_qual = seq, qual_full
        else:
            out_seq, out_qual = seq[s:e], qual_full[s:e]
            sample_idx = int(rec["sample"])
            if res in (ResolutionType.FULL_MATCH, ResolutionType.DEREPLICATED_FULL, ResolutionType.MULTIPLE_SPECIMENS):
                sample = tables.specimen_ids[sample_idx]
            elif res is ResolutionType.PARTIAL_FORWARD:
                sample = SampleId.PREFIX_FWD_MATCH + tables.b1[sample_idx]
            elif res is ResolutionType.PARTIAL_REVERSE:
                sample = SampleId.PREFIX_REV_MATCH + tables.b2[sample_idx]
            else:
                sample = SampleId.UNKNOWN
            pool = tables.pools[int(rec["pool"])] if rec["pool"] >= 0 else "unknown"
            p1n = tables.primer_names[int(rec["p1"])] if rec["p1"] >= 0 else "unknown"
            p2n = tables.primer_names[int(rec["p2"])] if rec["p2"] >= 0 else "unknown"
        ops.append(WriteOperation(
            sample_id=sample, seq_id=seq_ids[r], distance_code=code, sequence=out_seq,
            quality_sequence=out_qual, quality_scores=None,
            p1_location=_loc(rec["p1_loc"]), p2_location=_loc(rec["p2_loc"]),
            b1_location=_loc(rec["b1_loc"]), b2_location=_loc(rec["b2_loc"]),
            primer_pool=pool, p1_name=p1n, p2_name=p2n, resolution_type=res,
            trace_sequence_id=None if trace_ids is None else trace_ids[r]))
    return ops


def process_sequences(seq_records: List, parameters: MatchParameters, specimens, args: argparse.Namespace,
                      prefilter=None, trace_logger=None, record_offset: int = 0,
                      device: int = 0, _binding=None) -> Tuple[List[WriteOperation], int, int]:
    """reference: demultiplex.py:108-212 (same arguments; `device` selects the GPU)."""
    n = len(seq_records)
    if n == 0:
        return [], 0, 0
    matcher = get_matcher(parameters, specimens, args, prefilter, device, _binding)
    bases, quals, ids = [], [], []
    for rec in seq_records:
        b, q = get_bases_and_quality(rec)
        bases.append(b)
        quals.append(q)
        ids.append(rec.id)
    batch = PackedBatch(bases, clip=parameters.search_len)
    trace_ids = None
    if trace_logger is not None and getattr(trace_logger, "enabled", True):
        # tracing narrates every search: ask the device for the per-search detail arrays as well
        from .trace import emit_batch_trace
        result = matcher.match(batch, detail="locations" if getattr(trace_logger, "verbosity", 1) >= 3 else True)
        trace_ids = emit_batch_trace(trace_logger, matcher, result, seq_records, record_offset, args, specimens, parameters)
    else:
        result = matcher.match(batch, reuse=True)
    ops = records_to_write_ops(matcher.tables, result, ids, bases, quals, trace_ids)
    return ops, n, result.n_matched
