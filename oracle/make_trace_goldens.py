#!/usr/bin/env python3
"""oracle/make_trace_goldens.py -- TEST INFRASTRUCTURE.  Run in the BUILD container only (needs
/root/reference); writes tests/golden/trace_events.json.gz.

Runs the UNMODIFIED reference's process_sequences with the reference's own TraceLogger (over the same
stand-ins as make_goldens.py) on slices of the committed golden bundles, at -d1 / -d2 / -d3, and stores
the event streams (timestamps dropped).  tests/test_trace_events.py replays the same inputs through this
package (CPU kernel simulator / CUDA) and compares event by event.
"""
import csv
import glob
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import make_goldens as mg                      # noqa: E402  (sets up the reference + stand-ins)
from specimux.trace import TraceLogger as RefTraceLogger    # noqa: E402  (the real reference)

CASES = [  # (golden bundle, run, verbosity, reads)
    ("fixture", "default", 1, 40), ("fixture", "default", 2, 40), ("fixture", "default", 3, 40),
    ("fixture", "derep_none", 1, 40), ("fixture", "trim_tails", 1, 40), ("fixture", "no_preorient", 2, 40),
    ("fixture_rc", "default", 1, 40),
    ("fixture_crafted", "default", 1, 120), ("fixture_crafted", "no_prefilter", 3, 60),
    ("synth_ont037", "default", 1, 150), ("synth_ont037", "default", 3, 25), ("synth_ont037", "derep_none", 1, 80),
    ("synth_ont037_crafted", "no_prefilter", 2, 120),      # non-ACGT reads + big table: the reference side runs without the Bloom filter
    ("synth_multipool", "default", 1, 150), ("synth_multipool", "derep_none", 2, 60),
    ("synth_dense", "default", 1, 150), ("synth_dense", "derep_none", 1, 100), ("synth_dense", "trim_primers", 1, 60),
    ("synth_long", "default", 1, 40),
]


def load_golden(name):
    with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".json.gz"), "rt") as fh:
        return json.load(fh)


def reference_events(tmp, g, run, verbosity, n_reads):
    name = g["name"]
    ppath, spath = os.path.join(tmp, name + "_p.fasta"), os.path.join(tmp, name + "_s.txt")
    with open(ppath, "w") as fh:
        for pname, seq, pos, pools in g["primers"]:
            fh.write(">%s pool=%s position=%s\n%s\n" % (pname, ",".join(pools), pos, seq))
    with open(spath, "w") as fh:
        fh.write("SampleID\tPrimerPool\tFwIndex\tFwPrimer\tRvIndex\tRvPrimer\n")
        for row in g["specimens"]:
            fh.write("\t".join(row) + "\n")
    kw = dict(g["runs"][run]["flags"])
    big = len({r[2] for r in g["specimens"]} | {r[4] for r in g["specimens"]}) > 12
    if big:
        kw["disable_prefilter"] = True      # as make_goldens.bundle: result- and event-neutral on the reference side
    args = mg.make_args(**kw)
    registry = mg.ref_io.read_primers_file(ppath)
    specimens = mg.ref_io.read_specimen_file(spath, registry)
    specimens.validate()
    params = mg.ref_orch.setup_match_parameters(args, specimens)
    if args.disable_prefilter:
        prefilter = mg.PassthroughPrefilter()
    else:
        rcs = mg.barcodes_for_bloom_prefilter(specimens)
        prefilter = mg.BloomPrefilter.load_readonly(mg.BloomPrefilter.get_cache_path(rcs, params.max_dist_index),
                                                    rcs, params.max_dist_index)
    records = []
    for rid, s, q in g["reads"][:n_reads]:
        rec = mg.SeqRecord(mg.Seq(s), id=rid, name=rid, description=rid)
        if q is not None:
            rec.letter_annotations["phred_quality"] = [ord(c) - 33 for c in q]
        records.append(rec)
    out_dir = tempfile.mkdtemp(prefix="trace_", dir=tmp)
    tl = RefTraceLogger(True, verbosity, out_dir, "main", "T")
    mg.ref_demux.process_sequences(records, params, specimens, args, prefilter, tl, 0)
    tl.close()
    path = glob.glob(os.path.join(out_dir, "trace", "*.tsv"))[0]
    with open(path, newline="") as fh:
        rows = list(csv.reader(fh, delimiter="\t"))
    return [r[3:] for r in rows[1:]]        # sequence_id, event_type, fields...


def main():
    tmp = tempfile.mkdtemp(prefix="smx_trace_gold_")
    cases = []
    cache = {}
    for name, run, verbosity, n in CASES:
        g = cache.setdefault(name, load_golden(name))
        ev = reference_events(tmp, g, run, verbosity, n)
        cases.append(dict(golden=name, run=run, verbosity=verbosity, n_reads=n, events=ev))
        print("  %-22s %-14s -d%d %4d reads %6d events" % (name, run, verbosity, n, len(ev)))
    path = os.path.join(ROOT, "tests", "golden_trace", "trace_events.json.gz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with gzip.GzipFile(path, "wb", mtime=0) as fh:
        fh.write(json.dumps(dict(cases=cases), sort_keys=True).encode())
    print("wrote %s (%d KB)" % (path, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
