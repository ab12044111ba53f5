#!/usr/bin/env python3
"""oracle/make_goldens.py -- TEST INFRASTRUCTURE.  Run in the BUILD container only
(needs /root/reference); writes tests/golden/*.json.gz.

What it does, in order:
 1. Puts oracle/standins (edlib -> oracle/edlib_restated.c, Bio, pybloomfilter) and the
    UNMODIFIED reference package (/root/reference/src) on sys.path.  The only injected change is
    an insertion-ordered set for PrimerInfo.barcodes (the reference iterates a
    Python `set`, whose order depends on PYTHONHASHSEED; pinned here to first appearance).
 2. Runs the reference CLI on its own integration fixture and asserts that the output tree is
    byte-identical to tests/data/integration_test_suite/expected_output (this pins the aligner
    restatement against the reference's only golden vectors).
 3. For the fixture and for seeded synthetic slices of every BASELINE config (plus crafted
    edge cases), runs the reference's process_sequences and oracle.pipeline.process_reads on
    the same inputs under several flag sets, asserts equality of every write-operation, and
    stores the vectors.
 4. Cross-checks the exact-set Bloom emulation against the reference's BloomPrefilter and the
    C aligner against an independent pure-Python DP on random IUPAC strings.
"""
import argparse
import gzip
import hashlib
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
FIX = os.path.join(REF, "tests/data/integration_test_suite")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(HERE, "standins"))

os.environ["HOME"] = tempfile.mkdtemp(prefix="smx_home_")   # Bloom cache goes to ~/.specimux/cache

from oracle import pipeline as orc                          # noqa: E402
from oracle import aligner                                  # noqa: E402
from specimux_b200 import synth                             # noqa: E402

import specimux.models as ref_models                        # noqa: E402  (the real reference)


class OrderedSet:
    """Insertion-ordered stand-in for the `set` at models.py:28 (Q2 pin)."""

    def __init__(self):
        self._d = {}

    def add(self, x):
        self._d[x] = True

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __contains__(self, x):
        return x in self._d


_orig_init = ref_models.PrimerInfo.__init__


def _patched_init(self, *a, **kw):
    _orig_init(self, *a, **kw)
    self.barcodes = OrderedSet()


ref_models.PrimerInfo.__init__ = _patched_init

from specimux import io_utils as ref_io                     # noqa: E402
from specimux import orchestration as ref_orch              # noqa: E402
from specimux import demultiplex as ref_demux               # noqa: E402
from specimux.databases import PassthroughPrefilter         # noqa: E402
from specimux.bloom_filter import BloomPrefilter, barcodes_for_bloom_prefilter   # noqa: E402
from Bio.Seq import Seq                                     # noqa: E402
from Bio.SeqRecord import SeqRecord                         # noqa: E402

FLAG_DEFAULTS = dict(min_length=-1, max_length=-1, index_edit_distance=-1, primer_edit_distance=-1,
                     search_len=80, trim="barcodes", dereplicate="best", diagnostics=None, debug=False,
                     disable_prefilter=False, disable_preorient=False, output_to_files=True,
                     output_file_prefix="", output_dir=".", color=False, threads=1, isfastq=True)


def make_args(**kw):
    d = dict(FLAG_DEFAULTS)
    d.update(kw)
    return argparse.Namespace(**d)


def sha(s):
    return hashlib.sha1(s.encode()).hexdigest()[:16]


def loc(l):
    return None if l is None else [int(l[0]), int(l[1])]


def run_reference(workdir, primers_path, specimens_path, reads, args):
    """Reference process_sequences on in-memory reads; returns (ops, total, matched, k_idx, k_primers)."""
    registry = ref_io.read_primers_file(primers_path)
    specimens = ref_io.read_specimen_file(specimens_path, registry)
    specimens.validate()
    if args.disable_prefilter:
        params = ref_orch.setup_match_parameters(args, specimens)
        prefilter = PassthroughPrefilter()
    else:
        params = ref_orch.setup_match_parameters(args, specimens)
        rcs = barcodes_for_bloom_prefilter(specimens)
        prefilter = BloomPrefilter.load_readonly(BloomPrefilter.get_cache_path(rcs, params.max_dist_index),
                                                 rcs, params.max_dist_index)
    records = []
    for rid, s, q in reads:
        rec = SeqRecord(Seq(s), id=rid, name=rid, description=rid)
        if q is not None:
            rec.letter_annotations["phred_quality"] = [ord(c) - 33 for c in q]
        records.append(rec)
    ops, total, matched = ref_demux.process_sequences(records, params, specimens, args, prefilter, None, 0)
    out = []
    for op in ops:
        out.append(dict(sample_id=op.sample_id, seq_id=op.seq_id, distance_code=op.distance_code,
                        seq_sha=sha(op.sequence), seq_len=len(op.sequence), qual_sha=sha(op.quality_sequence),
                        p1=loc(op.p1_location), p2=loc(op.p2_location), b1=loc(op.b1_location),
                        b2=loc(op.b2_location), pool=op.primer_pool, p1_name=op.p1_name, p2_name=op.p2_name,
                        res=op.resolution_type.value))
    return out, total, matched, params.max_dist_index, dict(params.max_dist_primers)


def run_oracle(primer_records, specimen_rows, reads, args):
    tables = orc.Tables(primer_records, specimen_rows)
    params = orc.setup_params(tables, args.index_edit_distance, args.primer_edit_distance,
                              search_len=args.search_len, preorient=not args.disable_preorient,
                              prefilter=not args.disable_prefilter, trim=args.trim,
                              dereplicate=args.dereplicate, min_length=args.min_length,
                              max_length=args.max_length)
    ops, total, matched = orc.process_reads(tables, params, reads)
    out = []
    for op in ops:
        out.append(dict(sample_id=op.sample_id, seq_id=op.seq_id, distance_code=op.distance_code,
                        seq_sha=sha(op.sequence), seq_len=len(op.sequence), qual_sha=sha(op.quality_sequence),
                        p1=loc(op.p1_location), p2=loc(op.p2_location), b1=loc(op.b1_location),
                        b2=loc(op.b2_location), pool=op.primer_pool, p1_name=op.p1_name, p2_name=op.p2_name,
                        res=op.resolution_type))
    return out, total, matched, params.max_dist_index, dict(params.max_dist_primers)


def compare(tag, ref, mine):
    r_ops, r_tot, r_match, r_k, r_kp = ref
    m_ops, m_tot, m_match, m_k, m_kp = mine
    assert (r_k, r_kp) == (m_k, m_kp), "%s: thresholds differ %r vs %r" % (tag, (r_k, r_kp), (m_k, m_kp))
    assert (r_tot, r_match) == (m_tot, m_match), "%s: counts differ %r vs %r" % (tag, (r_tot, r_match), (m_tot, m_match))
    assert len(r_ops) == len(m_ops), "%s: %d vs %d ops" % (tag, len(r_ops), len(m_ops))
    for i, (a, b) in enumerate(zip(r_ops, m_ops)):
        assert a == b, "%s: op %d differs\n ref   %r\n oracle %r" % (tag, i, a, b)


FLAGSETS = {
    "default": {},
    "trim_primers": {"trim": "primers"},
    "trim_tails": {"trim": "tails"},
    "trim_none": {"trim": "none"},
    "derep_none": {"dereplicate": "none"},
    "no_prefilter": {"disable_prefilter": True},
    "no_preorient": {"disable_preorient": True},
    "derep_none_tails_nopre": {"dereplicate": "none", "trim": "tails", "disable_preorient": True},
}


def step_fixture_cli(tmp):
    """Step 2: reference CLI on its own fixture == its own expected_output, byte for byte."""
    out = os.path.join(tmp, "fixture_out")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "standins"), os.path.join(REF, "src")])
    env["PYTHONHASHSEED"] = "0"
    subprocess.check_call([sys.executable, "-m", "specimux.cli", os.path.join(FIX, "primers.fasta"),
                           os.path.join(FIX, "specimens.txt"), os.path.join(FIX, "sequences.fastq"),
                           "-F", "-O", out, "-d", "-t", "1"], env=env, stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    expected = {}
    n = 0
    for root, _dirs, files in os.walk(os.path.join(FIX, "expected_output")):
        for f in files:
            if f.endswith(".tsv"):
                continue
            rel = os.path.relpath(os.path.join(root, f), os.path.join(FIX, "expected_output"))
            exp = open(os.path.join(root, f)).read()
            got = open(os.path.join(out, rel)).read()
            assert exp == got, "fixture file differs: %s" % rel
            expected[rel] = exp
            n += 1
    produced = sum(len([f for f in fs if not f.endswith(".tsv") and f != "log.txt"])
                   for _r, _d, fs in os.walk(out))
    assert produced == n, "reference produced %d files, expected tree has %d" % (produced, n)
    print("fixture CLI run: %d files byte-identical to the reference's expected_output" % n)
    return expected


def bundle(name, primer_records, specimen_rows, reads, flagsets, tmp, extra=None, search_len=80, acgt_only=True):
    ppath = os.path.join(tmp, name + "_primers.fasta")
    spath = os.path.join(tmp, name + "_specimens.txt")
    with open(ppath, "w") as fh:
        for pname, seq, pos, pools in primer_records:
            fh.write(">%s pool=%s position=%s\n%s\n" % (pname, ",".join(pools), pos, seq))
    with open(spath, "w") as fh:
        fh.write("SampleID\tPrimerPool\tFwIndex\tFwPrimer\tRvIndex\tRvPrimer\n")
        for row in specimen_rows:
            fh.write("\t".join(row) + "\n")
    runs = {}
    for fname in flagsets:
        kw = dict(FLAGSETS[fname])
        kw["search_len"] = search_len
        # building the reference's Bloom cache is minutes of pure Python for ~100 barcodes at k=3;
        # large tables are run with the prefilter disabled on the reference side (result-neutral on
        # ACGT-only flanks) while the oracle is ALSO run with it enabled and must agree.
        big = len({r[2] for r in specimen_rows} | {r[4] for r in specimen_rows}) > 12
        ref_kw = dict(kw)
        if big:
            ref_kw["disable_prefilter"] = True
        ref = run_reference(tmp, ppath, spath, reads, make_args(**ref_kw))
        mine = run_oracle(primer_records, specimen_rows, reads, make_args(**kw))
        if big and not kw.get("disable_prefilter"):
            mine_off = run_oracle(primer_records, specimen_rows, reads, make_args(**ref_kw))
            compare("%s/%s (reference vs oracle, prefilter off)" % (name, fname), ref, mine_off)
            if acgt_only:
                compare("%s/%s (oracle prefilter on vs off)" % (name, fname), mine_off, mine)
        else:
            compare("%s/%s" % (name, fname), ref, mine)
        runs[fname] = dict(flags=kw, ops=mine[0], total=mine[1], matched=mine[2], k_index=mine[3],
                           k_primers=mine[4], reference_prefilter=not ref_kw.get("disable_prefilter", False))
        print("  %-12s %-24s %5d reads %5d ops  matched %d  OK" % (name, fname, mine[1], len(mine[0]), mine[2]))
    data = dict(name=name, primers=[list(p) for p in primer_records], specimens=[list(r) for r in specimen_rows],
                reads=[list(r) for r in reads], runs=runs, search_len=search_len)
    if extra:
        data.update(extra)
    path = os.path.join(ROOT, "tests", "golden", name + ".json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(data, sort_keys=True).encode())
    print("  wrote %s (%d KB)" % (path, os.path.getsize(path) // 1024))


def crafted_reads(ds_reads, rng):
    """Edge cases: short reads (Q1), N / IUPAC / lower-case / U symbols (Q6), empty and tiny reads."""
    out = []
    base = [r for r in ds_reads[:40]]
    for i, (rid, s, q) in enumerate(base):
        n = len(s)
        if i % 8 == 0:      # truncate to < search_len from the 3' end
            cut = rng.choice([0, 1, 5, 20, 46, 47, 60, 78, 79, 80, 81, 100])
            s, q = s[:cut], q[:cut]
        elif i % 8 == 1:    # truncate from the 5' end
            cut = rng.choice([30, 47, 66, 79, 80, 81, 120])
            s, q = s[n - cut:], q[n - cut:]
        elif i % 8 == 2:    # sprinkle N
            s = "".join("N" if rng.random() < 0.03 else c for c in s)
        elif i % 8 == 3:    # IUPAC codes in the read
            s = "".join(rng.choice("RYSWKMBDHV") if rng.random() < 0.03 else c for c in s)
        elif i % 8 == 4:    # lower-case run at both ends
            s = s[:60].lower() + s[60:]
        elif i % 8 == 5:    # U in the read
            s = "".join("U" if (c == "T" and rng.random() < 0.2) else c for c in s)
        elif i % 8 == 6:    # N only inside the barcode flank prefix region
            s = s[:3] + "N" + s[4:]
        out.append(("craft_%s" % rid, s, q))
    return out


def step_bloom_crosscheck():
    rng = random.Random(11)
    bcs = ["ACGTTGCA", "TTGACCAG", "GGATCCTA", "ANGTTGCA"]
    for k in (1, 2):
        bf = BloomPrefilter(bcs, k)
        for _ in range(3000):
            b = rng.choice(bcs)
            flank = "".join(rng.choice("ACGTN" if rng.random() < 0.2 else "ACGT") for _ in range(rng.randint(0, 12)))
            if rng.random() < 0.5:      # near-variant flank
                v = list(b)
                for _e in range(rng.randint(0, k + 1)):
                    p = rng.randrange(len(v) + 1)
                    op = rng.random()
                    if op < 0.4 and p < len(v):
                        v[p] = rng.choice("ACGT")
                    elif op < 0.7 and p < len(v):
                        del v[p]
                    else:
                        v.insert(p, rng.choice("ACGT"))
                flank = "".join(v) + flank
            assert bf.match(b, flank) == orc._bloom_yes(b, flank, k), (b, flank, k)
    print("bloom emulation == reference BloomPrefilter (exact set) on 6000 random cases")


def step_aligner_crosscheck():
    rng = random.Random(5)
    alpha = "ACGT" * 6 + "NRYSWKMBDHV" + "acgtnU"
    for it in range(4000):
        m = rng.randint(0, 24) if it % 50 else rng.choice([64, 65])
        n = rng.randint(0, 60)
        q = "".join(rng.choice(alpha) for _ in range(m))
        t = "".join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.6 and n > m > 0:
            p = rng.randrange(0, n - m + 1)
            t = t[:p] + q + t[p + m:]
            t = "".join(rng.choice("ACGT") if rng.random() < 0.1 else c for c in t)
        mode = rng.choice(["HW", "SHW", "NW"])
        k = rng.choice([-1, 0, 1, 3, 7])
        a = aligner.align(q, t, mode, "locations", k, aligner.IUPAC_PAIRS)
        b = aligner.align_py(q, t, mode, k)
        assert a["editDistance"] == b["editDistance"], (q, t, mode, k, a, b)
        if m and n and mode != "NW":
            assert a["locations"] == b["locations"], (q, t, mode, k, a, b)
    print("C aligner == independent Python DP on 4000 random IUPAC cases")


def main():
    only = set(sys.argv[1:])
    tmp = tempfile.mkdtemp(prefix="smx_gold_")
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    step_aligner_crosscheck()
    step_bloom_crosscheck()
    expected_tree = step_fixture_cli(tmp)

    primer_records = orc.read_primers_fasta(os.path.join(FIX, "primers.fasta"))
    specimen_rows = orc.read_specimens_tsv(os.path.join(FIX, "specimens.txt"))
    reads = orc.read_fastq(os.path.join(FIX, "sequences.fastq"))
    reads_rc = orc.read_fastq(os.path.join(FIX, "sequences_rc.fastq"))
    if not only or "fixture" in only:
        print("fixture (reference's own 40 reads, + the reverse-complemented copy)")
        bundle("fixture", primer_records, specimen_rows, reads, list(FLAGSETS), tmp,
               extra=dict(expected_output=expected_tree))
        bundle("fixture_rc", primer_records, specimen_rows, reads_rc, ["default", "trim_tails", "derep_none"], tmp)
        rng = random.Random(7)
        bundle("fixture_crafted", primer_records, specimen_rows,
               crafted_reads(reads, rng) + crafted_reads(reads_rc, rng), list(FLAGSETS), tmp, acgt_only=False)

    sizes = dict(ont037=400, multipool=400, dense=300, long=120)
    for cfg, n in sizes.items():
        if only and cfg not in only:
            continue
        print("synthetic %s" % cfg)
        ds = synth.CONFIGS[cfg](n_reads=n)
        rd = ds.reads()
        fl = ["default", "derep_none", "trim_tails", "trim_primers", "no_preorient"]
        bundle("synth_" + cfg, ds.primers, ds.specimens, rd, fl, tmp, search_len=ds.search_len)
    if only and "crafted" not in only:
        return
    ds = synth.ont037(n_reads=80)
    print("synthetic ont037 crafted edge cases")
    bundle("synth_ont037_crafted", ds.primers, ds.specimens, crafted_reads(ds.reads(), random.Random(9)),
           ["default", "derep_none", "trim_tails", "no_prefilter"], tmp, acgt_only=False)
    shutil.rmtree(tmp, ignore_errors=True)
    print("all goldens written")


if __name__ == "__main__":
    main()
