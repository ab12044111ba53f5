"""Diagnostic trace log (mirror of the reference's trace.py file format and event names).

Same TSV layout (timestamp, worker_id, event_seq, sequence_id, event_type, fields...), file naming
(trace/specimux_trace_<ts>_<worker>.tsv) and sequence ids (trace.py:96-114 of the reference).
The per-read pipeline runs on the GPU, so events are emitted per batch: the per-search detail arrays
(primer hits, barcode hits, explicit orientation flags) and the records come back from the device and
are narrated in the reference's event vocabulary and order: SEQUENCE_RECEIVED /
FILTERED, ORIENTATION_DETECTED, PRIMER_MATCHED, BARCODE_MATCHED, NO_MATCH_FOUND, MATCH_SCORED / DISCARDED,
DEREPLICATE_EXPANDED / SELECTED / PARTIAL_SELECTED / UNKNOWN_SELECTED, SPECIMEN_RESOLVED,
SEQUENCE_TRIM_EMPTY, SEQUENCE_OUTPUT, and at -d2/-d3 PRIMER_SEARCH / BARCODE_SEARCH.
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

    def log_orientation_detected(self, sequence_id, orientation, forward_score, reverse_score, confidence):
        self._log_event(sequence_id, "ORIENTATION_DETECTED", orientation, forward_score, reverse_score, f"{confidence:.3f}")

    # detailed events (reference trace.py:316-334): level 2 = successful primer searches, level 3 = everything
    def log_primer_search(self, sequence_id, primer_name, primer_direction, search_start, search_end, found,
                          edit_distance, match_position):
        if self.verbosity >= 2 and (self.verbosity >= 3 or found):
            self._log_event(sequence_id, "PRIMER_SEARCH", primer_name, primer_direction, search_start, search_end,
                            str(found).lower(), edit_distance, match_position)

    def log_barcode_search(self, sequence_id, barcode_name, barcode_type, primer_adjacent, search_start, search_end,
                           found, edit_distance, match_position):
        if self.verbosity >= 3:
            self._log_event(sequence_id, "BARCODE_SEARCH", barcode_name, barcode_type, primer_adjacent, search_start,
                            search_end, str(found).lower(), edit_distance, match_position)


# ---------------------------------------------------------------------------------------------
# Trace replay.  Every search result below comes from the GPU (smx_results detail arrays: primer
# hits, per-barcode hits, explicit orientation flags) and every output decision is the GPU's
# (smx_record); this code only DESCRIBES them in the reference's event vocabulary and order
# (reference demultiplex.py:108-212, 216-598, 668-820).  It never feeds anything back into the
# results: emit_batch_trace() checks that the decisions it narrates are the ones the records carry.


class _End:
    """One (strand, primer) search as a candidate sees it (demultiplex.py:748-820 match_one_end)."""
    __slots__ = ("primer", "matched", "dist", "match_pos", "match_end", "barcodes", "n_loc", "loc_starts", "tail", "loc_hits")

    def __init__(self, primer):
        self.primer = primer
        self.matched = False
        self.dist = -1
        self.match_pos = -1            # locations()[0] = (match_pos, match_end) in search coordinates
        self.match_end = -1
        self.tail = None               # (min start, max end) over every location of every barcode hit
        self.loc_hits = None           # -d3: per primer end location {barcode list position: distance}
        self.barcodes = []             # [(barcode, distance, search_start)] sorted by distance, stable
        self.n_loc = 0
        self.loc_starts = []           # barcode_search_start per primer end location


class _Cand:
    """CandidateMatch (models.py:72-95) by reference to its two ends."""
    __slots__ = ("cid", "e1", "e2", "pool", "order", "n", "shift")

    def __init__(self, cid, e1, e2, pool, order, n):
        self.cid, self.e1, self.e2, self.pool, self.order, self.n = cid, e1, e2, pool, order, n
        self.shift = 0                  # cumulative trim_locations() offset (models.py:322-328)

    def extent(self, mode, barcode_length):
        """Trim extents (models.py:278-320) in the candidate's current (possibly already shifted) coordinates.
        The forward primer / barcode were found on the other strand, so their locations are mirrored
        (AlignmentResult.reversed, models.py:57-63)."""
        n, sh = self.n, self.shift
        p1 = (n - self.e1.match_end - 1 - sh, n - self.e1.match_pos - 1 - sh) if self.e1.matched else None
        p2 = (self.e2.match_pos - sh, self.e2.match_end - sh) if self.e2.matched else None
        ps, pe = (p1[1] + 1 if p1 else 0), (p2[0] if p2 else n)
        if mode == "primers":
            return ps, pe
        if mode == "barcodes":
            return (p1[0] if p1 else 0), (p2[1] + 1 if p2 else n)
        if mode == "tails":
            s = e = -1
            if self.has_b1() and self.e1.tail is not None:
                s = n - self.e1.tail[1] - 1 - sh
            if self.has_b2() and self.e2.tail is not None:
                e = self.e2.tail[1] + 1 - sh
            if s == -1:
                s = max(0, ps - barcode_length)
            if e == -1:
                e = min(n, pe + barcode_length)
            return s, e
        return 0, n

    # --- the accessors trace.py:_extract_match_info and the selection code use
    def p1(self):
        return self.e1.primer if self.e1.matched else None

    def p2(self):
        return self.e2.primer if self.e2.matched else None

    def has_b1(self):
        return bool(self.e1.matched and self.e1.barcodes)

    def has_b2(self):
        return bool(self.e2.matched and self.e2.barcodes)

    @staticmethod
    def _best(end):
        if not (end.matched and end.barcodes):
            return []
        d0 = end.barcodes[0][1]
        return [b for b, d, _s in end.barcodes if abs(d - d0) < 1.0]        # models.py:116-126

    def best_b1(self):
        return self._best(self.e1)

    def best_b2(self):
        return self._best(self.e2)

    def b1_distance(self):
        return self.e1.barcodes[0][1] if self.has_b1() else -1

    def b2_distance(self):
        return self.e2.barcodes[0][1] if self.has_b2() else -1

    def p1_distance(self):
        return self.e1.dist if self.e1.matched else -1

    def p2_distance(self):
        return self.e2.dist if self.e2.matched else -1

    def full(self):
        return self.e1.matched and self.e2.matched and self.has_b1() and self.has_b2()

    def info(self):
        """trace.py:169-212 _extract_match_info."""
        p1, p2 = self.p1(), self.p2()
        b1s, b2s = self.best_b1(), self.best_b2()
        presence = ("both" if self.has_b1() and self.has_b2() else "forward_only" if self.has_b1()
                    else "reverse_only" if self.has_b2() else "none")
        total = sum(d for d in (self.p1_distance(), self.p2_distance(), self.b1_distance(), self.b2_distance()) if d >= 0)
        return (self.cid or "unknown", p1.name if p1 else "none", p2.name if p2 else "none",
                b1s[0] if b1s else "none", b2s[0] if b2s else "none", presence, total,
                self.p1_distance(), self.p2_distance(), self.b1_distance(), self.b2_distance())

    def score(self):                    # demultiplex.py:226-236
        p1, p2, b1, b2 = self.e1.matched, self.e2.matched, self.has_b1(), self.has_b2()
        if p1 and p2 and b1 and b2:
            return 5
        if p1 and p2 and (b1 or b2):
            return 4
        if (p1 or p2) and (b1 or b2):
            return 3
        if p1 and p2:
            return 2
        return 1 if (p1 or p2) else 0


class _Replay:
    """Per-batch view of the detail arrays."""

    def __init__(self, matcher, result, specimens, parameters, args):
        from .constants import Primer
        self.t = matcher.tables
        self.res = result
        self.specimens = specimens
        self.parameters = parameters
        self.args = args
        self.nP = self.t.n_primers
        self.index = {id(p): i for i, p in enumerate(self.t.primers)}
        self.fwd = specimens.get_primers(Primer.FWD)
        self.rev = specimens.get_primers(Primer.REV)
        self.total_list = self.t.pb_off[-1]
        self.L = parameters.search_len
        self.loc_hits = None
        if getattr(result, "barcode_loc_hits", None) is not None:
            self.loc_hits = {}
            for h in result.barcode_loc_hits:
                self.loc_hits.setdefault((int(h["read"]), int(h["slot"]), int(h["location"])), {})[int(h["hit"]["barcode"])] = \
                    int(h["hit"]["distance"])

    def end(self, r, strand, primer, n):
        """The search of `primer` on `strand` (0 = read as given, 1 = its reverse complement)."""
        p = self.index[id(primer)]
        ph = self.res.primer_hits[strand * self.nP + p, r]
        e = _End(primer)
        if int(ph["distance"]) < 0:
            return e
        e.matched, e.dist, e.match_pos, e.n_loc = True, int(ph["distance"]), int(ph["first_start"]), int(ph["n_locations"])
        e.match_end = int(ph["first_end"])
        base = self.total_list * strand + self.t.pb_off[p]
        hits = []
        for j, bc in enumerate(primer.barcodes):
            bh = self.res.barcode_hits[base + j, r]
            if int(bh["distance"]) >= 0:
                hits.append((bc, int(bh["distance"]), int(bh["search_start"])))
                # SHW locations of this barcode: (shift, shift + column) for every equal-best end column
                shift = 0 if int(bh["search_start"]) == -1 else int(bh["search_start"])
                mask = int(bh["end_mask"])
                lo, hi = shift, shift + mask.bit_length() - 1
                e.tail = (lo, hi) if e.tail is None else (min(e.tail[0], lo), max(e.tail[1], hi))
        hits.sort(key=lambda h: h[1])                      # add_barcode_match re-sorts by distance, stable
        e.barcodes = hits
        # barcode_search_start of every equal-best primer end (BARCODE_SEARCH events): end + 1 in
        # align_seq coordinates; the end mask is in staged-window coordinates
        woff = max(n - self.L, 0)
        raw = n - self.L
        delta = 0 if raw >= 0 or raw == -1 else raw - max(n + raw, 0)
        for w in range(self.res.endmask.shape[1]):
            word = int(self.res.endmask[strand * self.nP + p, w, r])
            while word:
                low = word & -word
                pos = 32 * w + low.bit_length() - 1
                e.loc_starts.append(woff + pos + delta + 1)
                word ^= low
        if self.loc_hits is not None:
            e.loc_hits = [self.loc_hits.get((r, strand * self.nP + p, li), {}) for li in range(len(e.loc_starts))]
        return e

    def orientation(self, r, n, flagged):
        """determine_orientation (demultiplex.py:602-638) from the GPU's search results."""
        from .constants import Orientation
        if not self.parameters.preorient:
            return Orientation.UNKNOWN, 0, 0
        irregular = (n - self.L) < -1 or flagged
        fwd = rev = 0
        for plist, is_fwd in ((self.fwd, True), (self.rev, False)):
            for primer in plist:
                p = self.index[id(primer)]
                if irregular:
                    hit_s = int(self.res.orient_hits[p, r])
                    hit_rs = int(self.res.orient_hits[self.nP + p, r])
                else:       # head-window test of a strand == tail-window match on the other strand
                    hit_s = int(self.res.primer_hits[self.nP + p, r]["distance"]) >= 0
                    hit_rs = int(self.res.primer_hits[p, r]["distance"]) >= 0
                if is_fwd:
                    fwd += hit_s
                    rev += hit_rs
                else:
                    fwd += hit_rs
                    rev += hit_s
        if fwd > 0 and rev == 0:
            return Orientation.FORWARD, fwd, rev
        if rev > 0 and fwd == 0:
            return Orientation.REVERSE, fwd, rev
        return Orientation.UNKNOWN, fwd, rev


def _log_end_searches(tl, sid, end, which_primer, which_barcode, n, L):
    """PRIMER_SEARCH / BARCODE_SEARCH events of one match_one_end call (levels 2 and 3; the logger
    filters by verbosity).  Level 3 reads the per-location barcode hits (smx_results.barcode_loc_hits)."""
    ss, se = n - L, n
    tl.log_primer_search(sid, end.primer.name, which_primer, ss, se, False, -1, -1)
    if not end.matched:
        tl.log_primer_search(sid, end.primer.name, which_primer, ss, se, False, -1, -1)
        return
    tl.log_primer_search(sid, end.primer.name, which_primer, ss, se, True, end.dist, end.match_pos)
    if tl.verbosity < 3:
        return
    won = {b: (d, s) for b, d, s in end.barcodes}
    for j, b in enumerate(end.primer.barcodes):
        for li, bs in enumerate(end.loc_starts):
            tl.log_barcode_search(sid, b, which_barcode, end.primer.name, bs, n, False, -1, -1)
            if end.loc_hits is not None:
                if j in end.loc_hits[li]:
                    tl.log_barcode_search(sid, b, which_barcode, end.primer.name, bs, n, True, end.loc_hits[li][j], bs)
            elif b in won and won[b][1] == bs:      # merged detail only: the location that won the barcode
                tl.log_barcode_search(sid, b, which_barcode, end.primer.name, bs, n, True, won[b][0], bs)


def _file_index(p, missing):
    return p.file_index if p is not None else missing


def _replay_read(rp, tl, sid, r, rec, n, flagged):
    """Events between SEQUENCE_RECEIVED and the write operations of one unfiltered read; returns the
    [(candidate or None, sample id or None, resolution name or None)] the reference would turn into
    write operations, in order."""
    from .constants import Orientation, ResolutionType, SampleId
    from .tables import pool_from_primers
    specimens, args = rp.specimens, rp.args
    orientation, fs, rs_ = rp.orientation(r, n, flagged)
    conf = abs(fs - rs_) / (fs + rs_) if rp.parameters.preorient and fs + rs_ > 0 else 0.0
    tl.log_orientation_detected(sid, orientation.to_string(), fs, rs_, conf)
    cands = []
    ends = {}

    def end_of(strand, primer):
        key = (strand, id(primer))
        if key not in ends:
            ends[key] = rp.end(r, strand, primer, n)
        return ends[key]

    for fwd in rp.fwd:
        for rev in specimens.get_paired_primers(fwd.primer):
            for rc, used in ((0, "as_is"), (1, "reverse_complement")):
                if rc == 0 and orientation is Orientation.REVERSE:
                    continue
                if rc == 1 and orientation is Orientation.FORWARD:
                    continue
                e1 = end_of(0 if rc else 1, fwd)        # as-is: forward primer searched on the RC strand
                e2 = end_of(1 if rc else 0, rev)
                if tl.verbosity >= 2:
                    _log_end_searches(tl, sid, e1, "forward", "forward", n, rp.L)
                    _log_end_searches(tl, sid, e2, "reverse", "reverse", n, rp.L)
                if e1.matched or e2.matched:
                    pool = pool_from_primers(fwd, rev)
                    c = _Cand("%s_match_%d" % (sid, len(cands)), e1, e2, pool, len(cands), n)
                    (cid, p1n, p2n, b1n, b2n, _pres, _tot, p1d, p2d, b1d, b2d) = c.info()
                    mtype = "both" if e1.matched and e2.matched else "forward_only" if e1.matched else "reverse_only"
                    tl._log_event(sid, "PRIMER_MATCHED", cid, mtype, p1n, p2n, p1d, p2d, pool or "none", used)
                    btype = ("both" if c.has_b1() and c.has_b2() else "forward_only" if c.has_b1()
                             else "reverse_only" if c.has_b2() else "none")
                    tl._log_event(sid, "BARCODE_MATCHED", cid, btype, b1n, b2n, b1d, b2d, p1n, p2n)
                    cands.append(c)
    if not cands:
        tl.log_no_match_found(sid, "primer_search", "No primer matches found")
        return [(None, SampleId.UNKNOWN, ResolutionType.UNKNOWN)]

    # select_best_matches (demultiplex.py:216-259)
    scored = [(c.score(), c) for c in cands]
    for sc, c in scored:
        (cid, p1n, p2n, b1n, b2n, pres, tot, *_rest) = c.info()
        tl._log_event(sid, "MATCH_SCORED", cid, p1n, p2n, b1n, b2n, tot, pres, "%.3f" % float(sc))
    scored.sort(key=lambda x: x[0], reverse=True)
    top = scored[0][0]
    best = [c for sc, c in scored if sc == top]
    for sc, c in scored:
        if sc < top:
            (cid, p1n, p2n, b1n, b2n, *_rest) = c.info()
            tl._log_event(sid, "MATCH_DISCARDED", cid, p1n, p2n, b1n, b2n, float(sc), "lower_score")

    def resolve(c):
        """resolve_specimen (demultiplex.py:541-598) incl. its SPECIMEN_RESOLVED event."""
        sample, pool = SampleId.UNKNOWN, c.pool
        if c.full():
            ids = specimens.specimens_for_barcodes_and_primers(c.best_b1(), c.best_b2(), c.p1(), c.p2())
            if len(ids) > 1:
                sample, res = ids[0], ResolutionType.MULTIPLE_SPECIMENS
                pool = specimens.get_specimen_pool(sample)
            elif len(ids) == 1:
                sample, res = ids[0], ResolutionType.FULL_MATCH
                pool = specimens.get_specimen_pool(sample)
            else:
                res = ResolutionType.UNKNOWN
        else:
            b1s, b2s = c.best_b1(), c.best_b2()
            if c.has_b1() and not c.has_b2() and len(b1s) == 1:
                sample, res = SampleId.PREFIX_FWD_MATCH + b1s[0], ResolutionType.PARTIAL_FORWARD
            elif c.has_b2() and not c.has_b1() and len(b2s) == 1:
                sample, res = SampleId.PREFIX_REV_MATCH + b2s[0], ResolutionType.PARTIAL_REVERSE
            else:
                res = ResolutionType.UNKNOWN
        c.pool = pool
        (_cid, p1n, p2n, b1n, b2n, *_rest) = c.info()
        tl._log_event(sid, "SPECIMEN_RESOLVED", sample, res.to_string(), pool or "none", p1n, p2n, b1n, b2n)
        return sample, res

    out = []
    if getattr(args, "dereplicate", "best") != "best":
        for c in best:
            sample, res = resolve(c)
            out.append((c, sample, res))
        return out

    # dereplicate_matches (demultiplex.py:262-393)
    expanded = []
    for c in best:
        if not c.full():
            expanded.append((c, None, None, None, 999.0, 999.0))
            continue
        d1 = {b: d for b, d, _s in c.e1.barcodes}
        d2 = {b: d for b, d, _s in c.e2.barcodes}
        found = False
        for b1 in c.best_b1():
            for b2 in c.best_b2():
                spec = specimens.specimen_for_exact_match(b1, b2, c.p1(), c.p2())
                if spec:
                    expanded.append((c, spec, b1, b2, d1.get(b1, 999.0), d2.get(b2, 999.0)))
                    found = True
        if not found:
            expanded.append((c, None, None, None, 999.0, 999.0))
    tl._log_event(sid, "DEREPLICATE_EXPANDED", len(best), len(expanded))
    groups = {}
    for e in expanded:
        groups.setdefault(e[1], []).append(e)
    results = []
    for spec, group in groups.items():
        if spec is None:
            members = [e[0] for e in group]
            single = [c for c in members if c.has_b1() != c.has_b2()]
            none = [c for c in members if not c.has_b1() and not c.has_b2()]
            other = [c for c in members if c.has_b1() and c.has_b2()]
            if single:                                  # dereplicate_partial_matches (:396-477)
                bgroups = {}
                for c in single:
                    direction = "forward" if c.has_b1() else "reverse"
                    for b in (c.best_b1() if c.has_b1() else c.best_b2()):
                        bgroups.setdefault((direction, b), []).append(c)
                for (direction, b), g in bgroups.items():
                    def key(c, direction=direction):
                        bd = c.b1_distance() if direction == "forward" else c.b2_distance()
                        cnt = (c.p1() is not None) + (c.p2() is not None)
                        pd = (c.p1_distance() if c.p1() else 0) + (c.p2_distance() if c.p2() else 0)
                        fi = _file_index(c.p1(), 0) + _file_index(c.p2(), 0)
                        return (bd, -cnt, pd, fi)
                    g = sorted(g, key=key)
                    k = key(g[0])
                    tl._log_event(sid, "DEREPLICATE_PARTIAL_SELECTED", direction, b, len(g), k[0], -k[1], k[2], k[3])
                    results.append((g[0], None))
            if none:                                    # dereplicate_unknown_matches (:480-538)
                def ukey(c):
                    cnt = (c.p1() is not None) + (c.p2() is not None)
                    pd = (c.p1_distance() if c.p1() else 0) + (c.p2_distance() if c.p2() else 0)
                    return (-cnt, pd, _file_index(c.p1(), 999) + _file_index(c.p2(), 999))
                g = sorted(none, key=ukey)
                if len(none) > 1:
                    k = ukey(g[0])
                    tl._log_event(sid, "DEREPLICATE_UNKNOWN_SELECTED", len(none), -k[0], k[1], k[2])
                results.append((g[0], None))
            for c in other:
                results.append((c, None))
            continue

        def skey(e):
            c = e[0]
            return (e[4] + e[5], c.p1_distance() + c.p2_distance(), _file_index(c.p1(), 999) + _file_index(c.p2(), 999))
        group = sorted(group, key=skey)
        k = skey(group[0])
        results.append((group[0][0], spec))
        tl._log_event(sid, "DEREPLICATE_SELECTED", spec, len(group), k[0], k[1], k[2])
    for c, spec in results:
        if spec is not None:
            c.pool = specimens.get_specimen_pool(spec)
            out.append((c, spec, ResolutionType.DEREPLICATED_FULL))
        else:
            sample, res = resolve(c)
            out.append((c, sample, res))
    return out


def emit_batch_trace(trace_logger: TraceLogger, matcher, result, seq_records, record_offset, args,
                     specimens=None, parameters=None):
    """Per-batch emission; returns the per-read trace sequence ids.  `result` must carry the detail arrays
    (Matcher.match(detail=True)).  Raises if the replayed decisions differ from the GPU's records."""
    tables = matcher.tables
    specimens = specimens if specimens is not None else tables.specimens
    parameters = parameters if parameters is not None else tables.parameters
    rp = _Replay(matcher, result, specimens, parameters, args)
    ids = []
    min_len, max_len = getattr(args, "min_length", -1), getattr(args, "max_length", -1)
    off = result.rec_offset
    acgt = set("ACGT")
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
        seq = rec.seq if isinstance(rec.seq, str) else str(rec.seq)
        flagged = not set(seq) <= acgt
        expected = _replay_read(rp, trace_logger, sid, i, rec, n, flagged)
        records = result.records[off[i]:off[i + 1]]
        if len(records) != len(expected):
            raise RuntimeError("trace replay: read %s yields %d write operations, the GPU returned %d"
                               % (rec.id, len(expected), len(records)))
        mode = getattr(args, "trim", "barcodes")
        for r, (c, sample, res) in zip(records, expected):
            s, e = c.extent(mode, specimens.b_length()) if c is not None else (0, n)
            if r["trim_empty"]:
                p1 = c.p1().name if c is not None and c.p1() else "unknown"
                p2 = c.p2().name if c is not None and c.p2() else "unknown"
                trace_logger.log_sequence_trim_empty(sid, mode, s, e, n, p1, p2)
                continue
            if int(r["resolution"]) != res.value:
                raise RuntimeError("trace replay: read %s resolves to %s, the GPU record says %d"
                                   % (rec.id, res, int(r["resolution"])))
            if mode != "none":
                if (s, e) != (int(r["trim_start"]), int(r["trim_end"])):
                    raise RuntimeError("trace replay: read %s trims to [%d, %d), the GPU record says [%d, %d)"
                                       % (rec.id, s, e, int(r["trim_start"]), int(r["trim_end"])))
                if c is not None:
                    c.shift += s        # create_write_operation mutates the candidate (demultiplex.py:76)
    return ids
