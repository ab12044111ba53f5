"""Shared test helpers: golden bundles, table construction through the product's own classes,
op comparison, and the (test-only) CPU kernel simulator binding."""
import argparse
import ctypes
import gzip
import hashlib
import json
import os
import subprocess

from specimux_b200 import _lib
from specimux_b200.constants import Primer
from specimux_b200.databases import BloomEmulationPrefilter, PassthroughPrefilter, PrimerDatabase, Specimens
from specimux_b200.models import MatchParameters, PrimerInfo
from specimux_b200.seqio import SeqRecord

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name + ".json.gz"), "rt") as fh:
        return json.load(fh)


def golden_names():
    return sorted(f[:-8] for f in os.listdir(GOLDEN) if f.endswith(".json.gz"))


def build_specimens(primer_records, specimen_rows):
    registry = PrimerDatabase()
    for idx, (name, seq, position, pools) in enumerate(primer_records):
        direction = Primer.FWD if position == "forward" else Primer.REV
        registry.add_primer(PrimerInfo(name, seq, direction, list(pools), file_index=idx), list(pools))
    registry.validate_pools()
    specimens = Specimens(registry)
    for sid, pool, b1, p1, b2, p2 in specimen_rows:
        specimens.add_specimen(sid, pool, b1.upper(), p1, b2.upper(), p2)
    specimens.validate()
    return specimens


def make_args(flags):
    d = dict(min_length=-1, max_length=-1, trim="barcodes", dereplicate="best", disable_prefilter=False,
             disable_preorient=False, search_len=80, output_to_files=True, isfastq=True, diagnostics=None)
    d.update(flags)
    return argparse.Namespace(**d)


def params_from_run(run, specimens):
    return MatchParameters(dict(run["k_primers"]), run["k_index"], run["flags"].get("search_len", 80),
                           not run["flags"].get("disable_preorient", False))


def prefilter_for(args):
    return PassthroughPrefilter() if args.disable_prefilter else BloomEmulationPrefilter()


def sha(s):
    return hashlib.sha1(s.encode()).hexdigest()[:16]


def op_to_dict(op):
    loc = lambda l: None if l is None else [int(l[0]), int(l[1])]
    res = op.resolution_type
    return dict(sample_id=op.sample_id, seq_id=op.seq_id, distance_code=op.distance_code,
                seq_sha=sha(op.sequence), seq_len=len(op.sequence), qual_sha=sha(op.quality_sequence),
                p1=loc(op.p1_location), p2=loc(op.p2_location), b1=loc(op.b1_location), b2=loc(op.b2_location),
                pool=op.primer_pool, p1_name=op.p1_name, p2_name=op.p2_name,
                res=res if isinstance(res, int) else res.value)


def records(reads):
    return [SeqRecord(s, rid, rid, q) for rid, s, q in reads]


def assert_ops_equal(got, expected, tag=""):
    assert len(got) == len(expected), "%s: %d ops vs %d expected" % (tag, len(got), len(expected))
    for i, (a, b) in enumerate(zip(got, expected)):
        assert a == b, "%s: op %d differs\n got      %r\n expected %r" % (tag, i, a, b)


_hostsim = None


def hostsim_binding():
    """Builds (g++) and loads tests/hostsim/libhostsim.so -- the CPU simulator of the CUDA kernels'
    per-thread routines.  TEST ONLY; the product never loads it."""
    global _hostsim
    if _hostsim is None:
        d = os.path.join(HERE, "hostsim")
        so = os.path.join(d, "libhostsim.so")
        srcs = [os.path.join(d, "hostsim.cpp")] + [os.path.join(ROOT, "specimux_b200", "csrc", f)
                                                   for f in ("smx_core.cuh", "smx_kernels.cuh", "smx_host_tables.hpp")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function", "-Wno-maybe-uninitialized",
                                   "-o", so, srcs[0]])
        lib = ctypes.CDLL(so)
        lib.hostsim_last_error.restype = ctypes.c_char_p
        lib.hostsim_match_batch.argtypes = [ctypes.POINTER(_lib.SmxTables), ctypes.POINTER(_lib.SmxParams),
                                            ctypes.POINTER(_lib.SmxBatch), ctypes.POINTER(_lib.SmxResults)]
        lib.hostsim_check_long_carry.argtypes = [ctypes.c_ulonglong]
        _hostsim = lib
    return _hostsim
