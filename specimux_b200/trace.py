"""Diagnostic trace log (mirror of the reference's trace.py file format and event names).

Same TSV layout (timestamp, worker_id, event_seq, sequence_id, event_type, fields...), file naming
(trace/specimux_trace_<ts>_<worker>.tsv) and sequence ids (trace.py:96-114 of the reference).
The per-read pipeline runs on the GPU, so events are emitted per batch from the returned records:
SEQUENCE_RECEIVED, SEQUENCE_FILTERED, NO_MATCH_FOUND, SPECIMEN_RESOLVED, DEREPLICATE_SELECTED,
SEQUENCE_TRIM_EMPTY and SEQUENCE_OUTPUT.  The per-candidate events (ORIENTATION_DETECTED,
PRIMER_MATCHED, BARCODE_MATCHED, MATCH_SCORED, MATCH_DISCARDED, PRIMER_SEARCH, BARCODE_SEARCH) need the
per-search detail arrays and are the next row of the scope table.
"""
import csv
from datetime import datetime
from pathlib import Path
from typing import Optional

from .constants import ResolutionType


class TraceLogger:
    def __init__(self, enabled: bool, verbosity: int, output_dir: str, worker_id: str, start_timestamp: str,
                 buffer_size: int = 1000):
        self.enabled = enabled
        self.verbosity = verbosity
        self.worker_id = worker_id
        self.event_counter = 0
        self.buffer = []
        self.buffer_size = buffer_size
        self.file_handle = None
        self.sequence_record_counter = 0
        if self.enabled:
            trace_dir = Path(output_dir) / "trace"
            trace_dir.mkdir(parents=True, exist_ok=True)
            self.filepath = trace_dir / f"specimux_trace_{start_timestamp}_{worker_id}.tsv"
            self.file_handle = open(self.filepath, "w", newline="")
            self.writer = csv.writer(self.file_handle, delimiter="\t")
            self.writer.writerow(["timestamp", "worker_id", "event_seq", "sequence_id", "event_type"])

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()

    def close(self):
        if self.enabled and self.file_handle:
            self._flush_buffer()
            self.file_handle.close()
            self.file_handle = None

    def _flush_buffer(self):
        if self.file_handle and self.buffer:
            self.writer.writerows(self.buffer)
            self.file_handle.flush()
            self.buffer = []

    def _log_event(self, sequence_id: str, event_type: str, *fields):
        if not self.enabled:
            return
        self.event_counter += 1
        self.buffer.append([datetime.now().isoformat(), self.worker_id, self.event_counter, sequence_id, event_type,
                            *fields])
        if len(self.buffer) >= self.buffer_size:
            self._flush_buffer()

    def get_sequence_id(self, seq_record, record_num: Optional[int] = None) -> str:
        if record_num is None:
            self.sequence_record_counter += 1
            record_num = self.sequence_record_counter
        return f"{seq_record.id}#{record_num:08d}#{self.worker_id}"

    def log_sequence_received(self, sequence_id, sequence_length, sequence_name):
        self._log_event(sequence_id, "SEQUENCE_RECEIVED", sequence_length, sequence_name)

    def log_sequence_filtered(self, sequence_id, sequence_length, filter_reason):
        self._log_event(sequence_id, "SEQUENCE_FILTERED", sequence_length, filter_reason)

    def log_no_match_found(self, sequence_id, stage_failed, reason):
        self._log_event(sequence_id, "NO_MATCH_FOUND", stage_failed, reason)

    def log_specimen_resolved(self, sequence_id, specimen_id, resolution_type, pool, p1_name, p2_name, b1_name, b2_name):
        self._log_event(sequence_id, "SPECIMEN_RESOLVED", specimen_id, resolution_type, pool, p1_name, p2_name,
                        b1_name, b2_name)

    def log_dereplicate_selected(self, sequence_id, specimen_id, alternatives_count, scores):
        self._log_event(sequence_id, "DEREPLICATE_SELECTED", specimen_id, alternatives_count, *scores)

    def log_sequence_trim_empty(self, sequence_id, trim_mode, trim_start, trim_end, seq_length, p1_name, p2_name):
        self._log_event(sequence_id, "SEQUENCE_TRIM_EMPTY", trim_mode, trim_start, trim_end, seq_length, p1_name, p2_name)

    def log_sequence_output(self, sequence_id, specimen_id, pool, primer_pair, file_path):
        self._log_event(sequence_id, "SEQUENCE_OUTPUT", specimen_id, pool, primer_pair, file_path)


def emit_batch_trace(trace_logger: TraceLogger, matcher, result, seq_records, record_offset, args):
    """Per-batch emission from the GPU records; returns the per-read trace sequence ids."""
    tables = matcher.tables
    ids = []
    min_len, max_len = getattr(args, "min_length", -1), getattr(args, "max_length", -1)
    off = result.rec_offset
    for i, rec in enumerate(seq_records):
        sid = trace_logger.get_sequence_id(rec, record_offset + i)
        ids.append(sid)
        n = len(rec)
        trace_logger.log_sequence_received(sid, n, rec.id)
        if min_len != -1 and n < min_len:
            trace_logger.log_sequence_filtered(sid, n, "too_short")
            continue
        if max_len != -1 and n > max_len:
            trace_logger.log_sequence_filtered(sid, n, "too_long")
            continue
        for r in result.records[off[i]:off[i + 1]]:
            res = ResolutionType(int(r["resolution"]))
            p1 = tables.primer_names[int(r["p1"])] if r["p1"] >= 0 else "none"
            p2 = tables.primer_names[int(r["p2"])] if r["p2"] >= 0 else "none"
            pool = tables.pools[int(r["pool"])] if r["pool"] >= 0 else "none"
            if r["p1"] < 0 and r["p2"] < 0 and not r["trim_empty"]:
                trace_logger.log_no_match_found(sid, "primer_search", "No primer matches found")
            elif res is ResolutionType.DEREPLICATED_FULL:
                d = r["dist"]
                trace_logger.log_dereplicate_selected(sid, tables.specimen_ids[int(r["sample"])], 1,
                                                      (int(d[1]) + int(d[2]), int(d[0]) + int(d[3]), ""))
            elif r["trim_empty"]:
                trace_logger.log_sequence_trim_empty(sid, getattr(args, "trim", "barcodes"), int(r["trim_start"]),
                                                     int(r["trim_end"]), n, p1, p2)
            else:
                sample = ("unknown" if res is ResolutionType.UNKNOWN else
                          tables.specimen_ids[int(r["sample"])] if res in (ResolutionType.FULL_MATCH, ResolutionType.MULTIPLE_SPECIMENS)
                          else ("barcode_fwd_" + tables.b1[int(r["sample"])]) if res is ResolutionType.PARTIAL_FORWARD
                          else ("barcode_rev_" + tables.b2[int(r["sample"])]))
                trace_logger.log_specimen_resolved(sid, sample, res.to_string(), pool, p1, p2, "", "")
    return ids
