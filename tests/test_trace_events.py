"""Trace emission: the event stream of process_sequences with a TraceLogger against the
UNMODIFIED reference's stream on the same reads (tests/golden_trace/trace_events.json.gz, written by
oracle/make_trace_goldens.py), at -d1 / -d2 / -d3, event by event and field by field.  Runs through the CPU
kernel simulator here and through CUDA in tests/test_gpu_parity.py."""
import csv
import glob
import gzip
import json
import os

import pytest

import helpers as H
from specimux_b200.demultiplex import process_sequences
from specimux_b200.trace import TraceLogger

_PATH = os.path.join(H.HERE, "golden_trace", "trace_events.json.gz")


def load_cases():
    with gzip.open(_PATH, "rt") as fh:
        return json.load(fh)["cases"]


def case_ids():
    return ["%s-%s-d%d" % (c["golden"], c["run"], c["verbosity"]) for c in load_cases()]


def product_events(case, tmp_path, binding):
    g = H.load_golden(case["golden"])
    run = g["runs"][case["run"]]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    args = H.make_args(run["flags"])
    params = H.params_from_run(run, specimens)
    reads = [tuple(r) for r in g["reads"]][:case["n_reads"]]
    tl = TraceLogger(True, case["verbosity"], str(tmp_path), "main", "T")
    ops, total, matched = process_sequences(H.records(reads), params, specimens, args, H.prefilter_for(args), tl, 0,
                                            _binding=binding)
    tl.close()
    path = glob.glob(os.path.join(str(tmp_path), "trace", "*.tsv"))[0]
    with open(path, newline="") as fh:
        rows = list(csv.reader(fh, delimiter="\t"))
    assert rows[0] == ["timestamp", "worker_id", "event_seq", "sequence_id", "event_type"]
    assert [int(r[2]) for r in rows[1:]] == list(range(1, len(rows)))            # event_seq counts up
    return [r[3:] for r in rows[1:]], ops


def compare(case, got):
    want = case["events"]
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, "event %d differs\n got      %r\n expected %r" % (i, a, b)
    assert len(got) == len(want)


@pytest.mark.parametrize("idx", range(len(load_cases())), ids=case_ids())
def test_trace_events_match_reference_stream(idx, tmp_path):
    case = load_cases()[idx]
    got, _ops = product_events(case, tmp_path, H.hostsim_binding())
    compare(case, got)
