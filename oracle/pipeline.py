"""oracle/pipeline.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of specimux's per-read matching path (reference v0.7.0), written
index-based rather than object-based.  Every function cites the reference
file:line it follows (paths relative to /root/reference/src/specimux/).

Parity status: PINNED.  oracle/make_goldens.py runs the *unmodified* reference
package over oracle/standins/ and checks (a) that run reproduces the reference's
own tests/data/integration_test_suite/expected_output byte for byte and (b) this
restatement produces identical write-operations on the fixture and on synthetic
slices of every BASELINE config; the resulting vectors live in tests/golden/.
Unpinned by any reference test (oracle-derived): HW start tie-break, SHW end
lists, pybloomfilter hash false positives (an exact set is used).

Set-order quirk: PrimerInfo.barcodes is a Python `set` in the
reference (models.py:28); here, and in the golden generator, barcode order is
pinned to first appearance in specimens.txt.
"""
from collections import OrderedDict, namedtuple

from . import aligner

FWD, REV = 0, 1                      # constants.py:100-103 (Primer.FWD / Primer.REV)
ORIENT_FWD, ORIENT_REV, ORIENT_UNK = 1, 2, 3   # constants.py:118-122

# ResolutionType values, constants.py:54-61
FULL_MATCH, PARTIAL_FORWARD, PARTIAL_REVERSE, MULTIPLE_SPECIMENS, UNKNOWN, DEREPLICATED_FULL = 1, 2, 3, 4, 5, 6
RES_DIR = {FULL_MATCH: "full", MULTIPLE_SPECIMENS: "full", DEREPLICATED_FULL: "full",
           PARTIAL_FORWARD: "partial", PARTIAL_REVERSE: "partial", UNKNOWN: "unknown"}

_COMP_SRC = "ACGTMRWSYKVHDBXN"
_COMP_DST = "TGCAKYWSRMBDHVXN"
_COMP = str.maketrans(_COMP_SRC + _COMP_SRC.lower() + "Uu", _COMP_DST + _COMP_DST.lower() + "Aa")


def revcomp(s):
    """Bio.Seq.reverse_complement on a str: ambiguous-DNA table, case kept, U->A, others unchanged."""
    return s.translate(_COMP)[::-1]


WriteOp = namedtuple("WriteOp", "sample_id seq_id distance_code sequence quality_sequence "
                                "p1_location p2_location b1_location b2_location "
                                "primer_pool p1_name p2_name resolution_type "
                                "trim_start trim_end is_rc trim_empty",
                     defaults=(None, None, None, None))
# trim_start / trim_end / is_rc / trim_empty are intermediate values of create_write_operation
# (demultiplex.py:39-43,47,74; match.sequence is seq or rseq, :703,724) exposed so that the record-level
# parity tests can compare them with smx_record without re-deriving them from the sliced strings.


class Primer:
    """models.py:20-31 (PrimerInfo); `barcodes` insertion-ordered (Q2 pinned)."""

    def __init__(self, name, seq, direction, pools, file_index):
        self.name = name
        self.primer = seq.upper()
        self.primer_rc = revcomp(self.primer)
        self.direction = direction
        self.pools = list(pools)
        self.file_index = file_index
        self.barcodes = OrderedDict()
        self.specimens = set()


class Tables:
    """io_utils.py:270-377 (parsers), databases.py:17-308 (PrimerDatabase, Specimens), restated."""

    def __init__(self, primer_records, specimen_rows):
        # primer_records: [(name, seq, 'forward'|'reverse', [pools])] in primers.fasta order
        self.registry = OrderedDict()            # name -> Primer        (databases.py:21)
        self.pool_primers = OrderedDict()        # pool -> {dir: [Primer]} (databases.py:23)
        for idx, (name, seq, position, pools) in enumerate(primer_records):
            if name in self.registry:
                raise ValueError("Duplicate primer name: %s" % name)
            p = Primer(name, seq, FWD if position == "forward" else REV, pools, idx)
            self.registry[name] = p
            for pool in pools:
                self.pool_primers.setdefault(pool, {FWD: [], REV: []})[p.direction].append(p)
        self.by_seq = OrderedDict()              # primer sequence -> Primer (databases.py:129,152-165; Q7)
        self.specimens = []                      # (id, pool, b1, [p1 objs], b2, [p2 objs]) (databases.py:167)
        self.barcode_length = 0
        active = set()
        seen = set()
        for sid, pool, b1, p1name, b2, p2name in specimen_rows:
            if sid in seen:
                raise ValueError("Duplicate specimen id in index file: %s" % sid)
            seen.add(sid)
            b1, b2 = b1.upper(), b2.upper()      # io_utils.py:344-345
            active.add(pool)
            self.barcode_length = max(self.barcode_length, len(b1), len(b2))
            p1s = self._resolve(p1name, pool, FWD)
            p2s = self._resolve(p2name, pool, REV)
            for plist, bc in ((p1s, b1), (p2s, b2)):
                for p in plist:
                    canon = self.by_seq.setdefault(p.primer, p)
                    canon.barcodes[bc] = True
                    canon.specimens.add(sid)
            self.specimens.append((sid, pool, b1, p1s, b2, p2s))
        # databases.py:169-195 prune_unused_pools (called from validate(), :281)
        for p in self.by_seq.values():
            p.pools = [x for x in p.pools if x in active]
        for pool in list(self.pool_primers):
            if pool not in active:
                del self.pool_primers[pool]
        self._pairs = {}
        # lookup indexes (same answers as the reference's linear scans, databases.py:219-245,266-271)
        self._rows_by_bc = {}
        self._pool_by_sid = {}
        for i, row in enumerate(self.specimens):
            self._rows_by_bc.setdefault((row[2], row[4]), []).append(i)
            self._pool_by_sid.setdefault(row[0], row[1])

    def _resolve(self, name, pool, direction):   # databases.py:197-217
        if name in ("-", "*"):
            d = self.pool_primers.get(pool, {FWD: [], REV: []})
            out = [p for p in d[FWD] + d[REV] if p.direction == direction]
            if not out:
                raise ValueError("No primers found in pool %s" % pool)
            return out
        p = self.registry.get(name)
        if p is None:
            raise ValueError("Primer not found: %s" % name)
        if p.direction != direction or pool not in self.pool_primers or p not in (
                self.pool_primers[pool][FWD] + self.pool_primers[pool][REV]):
            raise ValueError("Primer %s not usable in pool %s" % (name, pool))
        return [p]

    def primers(self, direction):                # databases.py:247-249
        return [p for p in self.by_seq.values() if p.direction == direction]

    def paired(self, primer):                    # databases.py:251-264
        if primer.primer not in self._pairs:
            self._pairs[primer.primer] = [q for q in self.by_seq.values()
                                          if q.direction != primer.direction and q.specimens & primer.specimens]
        return self._pairs[primer.primer]

    def specimen_pool(self, sid):                # databases.py:266-271 (first row with that id)
        return self._pool_by_sid.get(sid)

    def exact_specimen(self, b1, b2, p1, p2):    # databases.py:232-245 (identity test on primer objects;
        for i in self._rows_by_bc.get((b1, b2), ()):     # first row in file order, via the (b1, b2) index)
            sid, _pool, _sb1, p1s, _sb2, p2s = self.specimens[i]
            if any(p1 is x for x in p1s) and any(p2 is x for x in p2s):
                return sid
        return None

    def specimens_for(self, b1_list, b2_list, p1, p2):   # databases.py:219-230 (file order)
        rows = sorted({i for b1 in b1_list for b2 in b2_list for i in self._rows_by_bc.get((b1, b2), ())})
        out = []
        for i in rows:
            sid, _pool, _sb1, p1s, _sb2, p2s = self.specimens[i]
            if any(p1 is x for x in p1s) and any(p2 is x for x in p2s):
                out.append(sid)
        return out


class Params:
    """models.py:331-338 MatchParameters + the flags of cli.py:23-41 that change results."""

    def __init__(self, max_dist_primers, max_dist_index, search_len=80, preorient=True, prefilter=True,
                 trim="barcodes", dereplicate="best", min_length=-1, max_length=-1):
        self.max_dist_primers = max_dist_primers   # keyed by primer sequence (orchestration.py:607,612)
        self.max_dist_index = max_dist_index
        self.search_len = search_len
        self.preorient = preorient
        self.prefilter = prefilter
        self.trim = trim
        self.dereplicate = dereplicate
        self.min_length = min_length
        self.max_length = max_length


def nw_distance(a, b):
    return aligner.align(a, b, "NW", "distance", -1, None)["editDistance"]


def setup_params(tables, index_edit_distance=-1, primer_edit_distance=-1, **kw):
    """orchestration.py:548-641 setup_match_parameters (threshold derivation only)."""
    import itertools
    import math
    b1s, b2s = OrderedDict(), OrderedDict()
    for p in tables.primers(FWD):
        b1s.update(p.barcodes)
    for p in tables.primers(REV):
        b2s.update(p.barcodes)
    combined = list(b1s) + [revcomp(b) for b in b2s]          # :581
    if index_edit_distance != -1:
        k_idx = index_edit_distance
    else:
        # With <2 barcodes the reference's _sanity_check_distance returns None and :594 raises.
        dmin = min(nw_distance(x, y) for x, y in itertools.combinations(combined, 2))
        k_idx = math.ceil(dmin / 2.0)                          # :594

    def bp_adjusted(primer):                                   # :564-570
        score = 0
        for ch in primer:
            score += 3 if ch in "ACGT" else 2 if ch in "KMRSWY" else 1 if ch in "BDHV" else 0
        return score / 3.0

    thr = {}
    for p in tables.primers(FWD) + tables.primers(REV):        # :603-614
        thr[p.primer] = primer_edit_distance if primer_edit_distance != -1 else int(bp_adjusted(p.primer) / 3)
    return Params(thr, k_idx, **kw)


# ---------------------------------------------------------------------------------------------
# alignment primitive

def align_window(query, target, k, start, end, mode):
    """alignment.py:21-50 align_seq: Python-slice window (Q1), k clamp (Q4), shift by raw `s`.

    Returns (distance or -1, [(start, end), ...] in target coordinates)."""
    s = 0 if start == -1 else start
    e = len(target) if end == -1 else min(end, len(target))
    r = aligner.align(query, target[s:e], mode, "locations", k, aligner.IUPAC_PAIRS)
    d = r["editDistance"]
    if d != -1 and d > k:
        d = -1
    if d == -1:
        return -1, []
    return d, [(a + s, b + s) for a, b in r["locations"]]


def _bloom_yes(barcode_rc, flank, k):
    return aligner.bloom_yes(barcode_rc, flank, k)


def _bloom_yes_py(barcode_rc, flank, k):
    """bloom_filter.py:70-101,176-186 with an exact set (no hash false positives):
    key barcode_rc + flank[:m-k] is present iff some string within k edits of barcode_rc
    (edits introduce only A/C/G/T) has flank[:m-k] as its first m-k characters."""
    m = len(barcode_rc)
    t = flank[:m - k]
    if len(t) < m - k:
        return False
    if m - k <= 0:
        return True
    # constrained DP: rows over t, columns over prefixes of barcode_rc
    INF = 10 ** 6
    n = len(t)
    prev = list(range(m + 1))                  # t empty vs barcode prefix j: delete j chars
    for i in range(1, n + 1):
        ch = t[i - 1]
        creatable = ch in "ACGT"
        cur = [prev[0] + 1 if creatable else INF] + [0] * m
        for j in range(1, m + 1):
            best = prev[j - 1] if ch == barcode_rc[j - 1] else (prev[j - 1] + 1 if creatable else INF)
            if creatable:
                best = min(best, prev[j] + 1)   # ch inserted
            best = min(best, cur[j - 1] + 1)    # barcode char deleted
            cur[j] = min(best, INF)
        prev = cur
    return min(prev) <= k


def reverse_locs(locs, length):
    """models.py:52-63 AlignmentResult.reversed/_reverse: order kept."""
    return [(length - b - 1, length - a - 1) for a, b in locs]


def match_one_end(tables, params, X, primer):
    """demultiplex.py:748-820 on one strand `X` for one primer, in X coordinates.

    Returns (pdist, plocs, [(barcode, dist, locs), ...]) with barcodes in pinned order; the
    per-barcode result is the strictly-smallest distance over the primer's end locations
    (first location wins ties, :809-812)."""
    L = params.search_len
    pd, plocs = align_window(primer.primer_rc, X, params.max_dist_primers[primer.primer],
                             len(X) - L, len(X), "HW")
    hits = []
    if pd == -1:
        return -1, [], hits
    for b in primer.barcodes:
        b_rc = revcomp(b)
        best = None
        for (_ls, le) in plocs:
            bstart = le + 1
            flank = X[bstart:]                                   # :789 (Python slice, may be negative)
            if params.prefilter and not _bloom_yes(b_rc, flank, params.max_dist_index):
                continue
            bd, blocs = align_window(b_rc, X, params.max_dist_index, bstart, len(X), "SHW")
            if bd != -1 and (best is None or bd < best[0]):
                best = (bd, blocs)
        if best is not None:
            hits.append((b, best[0], best[1]))
    return pd, plocs, hits


def determine_orientation(tables, params, s, rs):
    """demultiplex.py:602-638."""
    fwd = rev = 0
    L = params.search_len
    for p in tables.primers(FWD):
        k = params.max_dist_primers[p.primer]
        fwd += align_window(p.primer, s, k, 0, L, "HW")[0] != -1
        rev += align_window(p.primer, rs, k, 0, L, "HW")[0] != -1
    for p in tables.primers(REV):
        k = params.max_dist_primers[p.primer]
        fwd += align_window(p.primer, rs, k, 0, L, "HW")[0] != -1
        rev += align_window(p.primer, s, k, 0, L, "HW")[0] != -1
    if fwd > 0 and rev == 0:
        return ORIENT_FWD, fwd, rev
    if rev > 0 and fwd == 0:
        return ORIENT_REV, fwd, rev
    return ORIENT_UNK, fwd, rev


def pool_from_primers(p1, p2):
    """demultiplex.py:640-665."""
    if p1 is not None and p2 is not None:
        common = set(p1.pools) & set(p2.pools)
        return sorted(common)[0] if common else None
    one = p1 if p1 is not None else p2
    if one is not None:
        return sorted(one.pools)[0] if one.pools else None
    return None


class Candidate:
    """models.py:72-328 CandidateMatch, holding coordinates in the candidate's own orientation."""

    def __init__(self, is_rc, length, barcode_length):
        self.is_rc = is_rc
        self.length = length
        self.blen = barcode_length
        self.p1 = self.p2 = None
        self.p1_dist = self.p2_dist = -1
        self.p1_locs = self.p2_locs = None
        self.b1 = []          # [barcode, dist, locs] sorted stably by dist (models.py:97-108)
        self.b2 = []
        self.pool = None

    def set_end(self, which, primer, result, rev):
        pd, plocs, hits = result
        if pd == -1:
            return
        plocs = reverse_locs(plocs, self.length) if rev else list(plocs)
        entries = []
        for b, bd, blocs in hits:
            entries.append([b, bd, reverse_locs(blocs, self.length) if rev else list(blocs)])
        entries.sort(key=lambda x: x[1])
        if which == FWD:
            self.p1, self.p1_dist, self.p1_locs, self.b1 = primer, pd, plocs, entries
        else:
            self.p2, self.p2_dist, self.p2_locs, self.b2 = primer, pd, plocs, entries

    def b1_dist(self):
        return self.b1[0][1] if self.b1 else -1

    def b2_dist(self):
        return self.b2[0][1] if self.b2 else -1

    def best_b1(self):                                   # models.py:116-126 (tolerance 1.0 on ints)
        return [e[0] for e in self.b1 if e[1] == self.b1[0][1]]

    def best_b2(self):
        return [e[0] for e in self.b2 if e[1] == self.b2[0][1]]

    def full(self):
        return self.p1 is not None and self.p2 is not None and bool(self.b1) and bool(self.b2)

    def distance_code(self):                             # models.py:206-218
        f = lambda d: str(d) if d >= 0 else "X"
        return "%s,%s,%s,%s" % (f(self.p1_dist), f(self.b1_dist()), f(self.b2_dist()), f(self.p2_dist))

    def extent(self, mode):                              # models.py:278-319
        ps, pe = 0, self.length
        if self.p1 is not None:
            ps = self.p1_locs[0][1] + 1
        if self.p2 is not None:
            pe = self.p2_locs[0][0]
        if mode == "primers":
            return ps, pe
        if mode == "barcodes":
            s, e = 0, self.length
            if self.p1 is not None:
                s = self.p1_locs[0][0]
            if self.p2 is not None:
                e = self.p2_locs[0][1] + 1
            return s, e
        s = e = -1                                        # tails, with the reference's -1 sentinel
        for ent in self.b1:
            for l in ent[2]:
                s = l[0] if s == -1 else min(s, l[0])
        for ent in self.b2:
            for l in ent[2]:
                e = l[1] + 1 if e == -1 else max(e, l[1] + 1)
        if s == -1:
            s = max(0, ps - self.blen)
        if e == -1:
            e = min(self.length, pe + self.blen)
        return s, e

    def shift(self, start):                              # models.py:321-328 trim_locations (mutating, Q3)
        mv = lambda locs: [(a - start, b - start) for a, b in locs]
        for ent in self.b1:
            ent[2] = mv(ent[2])
        for ent in self.b2:
            ent[2] = mv(ent[2])
        if self.p1 is not None:
            self.p1_locs = mv(self.p1_locs)
        if self.p2 is not None:
            self.p2_locs = mv(self.p2_locs)


def find_candidates(tables, params, s, rs, slot_cache=None):
    """demultiplex.py:668-746.  `slot_cache` memoises match_one_end per (primer, strand): the
    reference recomputes identical searches for every partner primer."""
    if slot_cache is None:
        slot_cache = {}
    if params.preorient:
        orient = determine_orientation(tables, params, s, rs)[0]
    else:
        orient = ORIENT_UNK
    strands = (s, rs)

    def slot(primer, strand):
        key = (primer.primer, strand)
        if key not in slot_cache:
            slot_cache[key] = match_one_end(tables, params, strands[strand], primer)
        return slot_cache[key]

    out = []
    n = len(s)
    for fwd in tables.primers(FWD):
        for rev in tables.paired(fwd):
            if orient in (ORIENT_FWD, ORIENT_UNK):
                c = Candidate(False, n, tables.barcode_length)
                c.set_end(FWD, fwd, slot(fwd, 1), True)
                c.set_end(REV, rev, slot(rev, 0), False)
                if c.p1 is not None or c.p2 is not None:
                    c.pool = pool_from_primers(fwd, rev)
                    out.append(c)
            if orient in (ORIENT_REV, ORIENT_UNK):
                c = Candidate(True, n, tables.barcode_length)
                c.set_end(FWD, fwd, slot(fwd, 0), True)
                c.set_end(REV, rev, slot(rev, 1), False)
                if c.p1 is not None or c.p2 is not None:
                    c.pool = pool_from_primers(fwd, rev)
                    out.append(c)
    return out


def score(c):
    """demultiplex.py:226-236."""
    both_p = c.p1 is not None and c.p2 is not None
    any_p = c.p1 is not None or c.p2 is not None
    any_b = bool(c.b1) or bool(c.b2)
    if both_p and c.b1 and c.b2:
        return 5
    if both_p and any_b:
        return 4
    if any_p and any_b:
        return 3
    if both_p:
        return 2
    return 1 if any_p else 0


def select_best(cands):
    """demultiplex.py:216-259: all candidates with the top score, original order."""
    scores = [score(c) for c in cands]
    top = max(scores)
    return [c for c, sc in zip(cands, scores) if sc == top]


def _pcount_pdist_fidx(c, missing):
    cnt = (c.p1 is not None) + (c.p2 is not None)
    dist = (c.p1_dist if c.p1 is not None else 0) + (c.p2_dist if c.p2 is not None else 0)
    fidx = (c.p1.file_index if c.p1 is not None else missing) + (c.p2.file_index if c.p2 is not None else missing)
    return cnt, dist, fidx


def dereplicate(tables, cands):
    """demultiplex.py:262-538 -> [(candidate, specimen_id or None)] in dict-insertion order."""
    expanded = []
    for c in cands:
        if not c.full():
            expanded.append((c, None, 999.0, 999.0))
            continue
        d1 = {e[0]: e[1] for e in c.b1}
        d2 = {e[0]: e[1] for e in c.b2}
        found = False
        for b1 in c.best_b1():
            for b2 in c.best_b2():
                sid = tables.exact_specimen(b1, b2, c.p1, c.p2)
                if sid:
                    expanded.append((c, sid, d1[b1], d2[b2]))
                    found = True
        if not found:
            expanded.append((c, None, 999.0, 999.0))
    groups = OrderedDict()
    for ent in expanded:
        groups.setdefault(ent[1], []).append(ent)
    results = []
    for sid, group in groups.items():
        if sid is None:
            single = [e[0] for e in group if bool(e[0].b1) != bool(e[0].b2)]
            none = [e[0] for e in group if not e[0].b1 and not e[0].b2]
            other = [e[0] for e in group if e[0].b1 and e[0].b2]
            if single:                                            # :396-477
                bgroups = OrderedDict()
                for c in single:
                    direction, bcs = ("forward", c.best_b1()) if c.b1 else ("reverse", c.best_b2())
                    for b in bcs:
                        bgroups.setdefault((direction, b), []).append(c)
                for (direction, _b), g in bgroups.items():
                    def key(c, direction=direction):
                        cnt, dist, fidx = _pcount_pdist_fidx(c, 0)
                        return (c.b1_dist() if direction == "forward" else c.b2_dist(), -cnt, dist, fidx)
                    results.append((sorted(g, key=key)[0], None))
            if none:                                              # :480-538
                def ukey(c):
                    cnt, dist, fidx = _pcount_pdist_fidx(c, 999)
                    return (-cnt, dist, fidx)
                results.append((sorted(none, key=ukey)[0], None))
            for c in other:
                results.append((c, None))
            continue

        def fkey(e):                                              # :371-378
            c = e[0]
            return (e[2] + e[3], c.p1_dist + c.p2_dist, c.p1.file_index + c.p2.file_index)
        results.append((sorted(group, key=fkey)[0][0], sid))
    return results


def resolve(tables, c):
    """demultiplex.py:541-598 -> (sample_id, resolution_type); may override c.pool."""
    if c.full():
        ids = tables.specimens_for(c.best_b1(), c.best_b2(), c.p1, c.p2)
        if len(ids) > 1:
            c.pool = tables.specimen_pool(ids[0])
            return ids[0], MULTIPLE_SPECIMENS
        if len(ids) == 1:
            c.pool = tables.specimen_pool(ids[0])
            return ids[0], FULL_MATCH
        return "unknown", UNKNOWN
    b1s, b2s = (c.best_b1() if c.b1 else []), (c.best_b2() if c.b2 else [])
    if c.b1 and not c.b2 and len(b1s) == 1:
        return "barcode_fwd_" + b1s[0], PARTIAL_FORWARD
    if c.b2 and not c.b1 and len(b2s) == 1:
        return "barcode_rev_" + b2s[0], PARTIAL_REVERSE
    return "unknown", UNKNOWN


def make_write_op(params, sample_id, seq_id, bases, quals, c, res):
    """demultiplex.py:30-103 create_write_operation; `quals` is the ASCII string or None (FASTA)."""
    s, e = 0, len(bases)
    first = lambda locs: locs[0] if locs else None
    locs = lambda: (first(c.p1_locs) if c.p1 is not None else None, first(c.p2_locs) if c.p2 is not None else None,
                    c.b1[0][2][0] if c.b1 else None, c.b2[0][2][0] if c.b2 else None)
    if params.trim != "none":
        s, e = c.extent(params.trim)
        if s >= e:                                                # :47-73 empty-trim fallback
            p1l, p2l, b1l, b2l = locs()
            return WriteOp("unknown", seq_id, c.distance_code(), bases, quals, p1l, p2l, b1l, b2l,
                           "unknown", "unknown", "unknown", UNKNOWN, 0, len(bases), c.is_rc, True)
        bases = bases[s:e]
        quals = quals[s:e] if quals is not None else None
        c.shift(s)
    p1l, p2l, b1l, b2l = locs()
    return WriteOp(sample_id, seq_id, c.distance_code(), bases, quals, p1l, p2l, b1l, b2l,
                   c.pool if c.pool else "unknown",
                   c.p1.name if c.p1 is not None else "unknown",
                   c.p2.name if c.p2 is not None else "unknown", res, s, e, c.is_rc, False)


def process_read(tables, params, seq_id, bases, quals):
    """demultiplex.py:126-210 for one read -> ([WriteOp], has_full_match)."""
    n = len(bases)
    if params.min_length != -1 and n < params.min_length:
        return [], False
    if params.max_length != -1 and n > params.max_length:
        return [], False
    rbases = revcomp(bases)
    rquals = quals[::-1] if quals is not None else None
    cands = find_candidates(tables, params, bases, rbases)
    ops = []
    full = False
    pick = lambda c: (rbases, rquals) if c.is_rc else (bases, quals)
    if not cands:
        c = Candidate(False, n, tables.barcode_length)
        ops.append(make_write_op(params, "unknown", seq_id, bases, quals, c, UNKNOWN))
        return ops, False
    best = select_best(cands)
    if params.dereplicate == "best":
        for c, sid in dereplicate(tables, best):
            sb, sq = pick(c)
            if sid is not None:
                c.pool = tables.specimen_pool(sid)
                ops.append(make_write_op(params, sid, seq_id, sb, sq, c, DEREPLICATED_FULL))
                full = True
            else:
                sample, res = resolve(tables, c)
                ops.append(make_write_op(params, sample, seq_id, sb, sq, c, res))
                full = full or res in (FULL_MATCH, DEREPLICATED_FULL)
    else:
        for c in best:
            sample, res = resolve(tables, c)
            sb, sq = pick(c)
            ops.append(make_write_op(params, sample, seq_id, sb, sq, c, res))
            full = full or res in (FULL_MATCH, DEREPLICATED_FULL)
    return ops, full


def process_reads(tables, params, reads):
    """demultiplex.py:108-212 process_sequences -> (ops, total, matched); reads = [(id, bases, quals)]."""
    ops, matched = [], 0
    for seq_id, bases, quals in reads:
        o, full = process_read(tables, params, seq_id, bases, quals)
        ops.extend(o)
        matched += bool(full)
    return ops, len(reads), matched


# ---------------------------------------------------------------------------------------------
# output contract (io_utils.py:197-268) -- used to compare byte-identical per-specimen files

def op_paths(op, prefix="", is_fastq=True):
    ext = ".fastq" if is_fastq else ".fasta"
    safe = "".join(ch if ch.isalnum() or ch in "._-$#" else "_" for ch in (op.sample_id or "unknown"))
    pool, p1, p2 = op.primer_pool or "unknown", op.p1_name or "unknown", op.p2_name or "unknown"
    top = RES_DIR[op.resolution_type]
    paths = ["%s/%s/%s-%s/%s%s%s" % (top, pool, p1, p2, prefix, safe, ext)]
    if top == "full":
        paths.append("full/%s/%s%s%s" % (op.primer_pool, prefix, safe, ext))
    return paths


def op_record(op, is_fastq=True):
    head = "%s %s pool=%s primers=%s+%s %s" % (op.seq_id, op.distance_code, op.primer_pool,
                                                op.p1_name, op.p2_name, op.sample_id)
    if is_fastq:
        return "@%s\n%s\n+\n%s\n" % (head, op.sequence, op.quality_sequence)
    return ">%s\n%s\n" % (head, op.sequence)


def render_tree(ops, prefix="", is_fastq=True):
    """{relative path: file content} for a list of write-ops in order (== `-t 1` output)."""
    tree = OrderedDict()
    for op in ops:
        rec = op_record(op, is_fastq)
        for p in op_paths(op, prefix, is_fastq):
            tree[p] = tree.get(p, "") + rec
    return tree


# ---------------------------------------------------------------------------------------------
# file parsers (io_utils.py:270-377), minimal

def read_primers_fasta(path):
    recs, title, chunks = [], None, []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if title is not None:
                    recs.append((title, "".join(chunks)))
                title, chunks = line[1:].rstrip(), []
            elif title is not None:
                chunks.append(line.strip())
    if title is not None:
        recs.append((title, "".join(chunks)))
    out = []
    for title, seq in recs:
        name = title.split()[0]
        pools, position = [], None
        for field in title.split():
            if field.startswith("pool="):
                pools = [p.strip() for p in field[5:].replace(";", ",").split(",")]
            elif field.startswith("position="):
                position = field[9:]
        out.append((name, seq, position, pools))
    return out


def read_specimens_tsv(path):
    import csv
    rows = []
    with open(path, newline="") as fh:
        for row in csv.DictReader(fh, delimiter="\t"):
            rows.append((row["SampleID"], row["PrimerPool"], row["FwIndex"], row["FwPrimer"],
                         row["RvIndex"], row["RvPrimer"]))
    return rows


def read_fastq(path):
    out = []
    with open(path) as fh:
        while True:
            t = fh.readline()
            if not t:
                break
            if not t.strip():
                continue
            s = fh.readline().strip()
            fh.readline()
            q = fh.readline().strip()
            out.append((t[1:].split()[0], s, q))
    return out
