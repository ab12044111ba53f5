"""CPU tests of the host layer: C-ABI exports, packer, table building, loud failure without a GPU,
file parsers, output tree rules (checked byte-for-byte against the reference's expected_output)."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers as H
from specimux_b200 import _lib
from specimux_b200.demultiplex import process_sequences
from specimux_b200.engine import Matcher, PackedBatch
from specimux_b200.io_utils import OutputManager, read_primers_file, read_specimen_file
from specimux_b200.seqio import parse, reverse_complement
from specimux_b200.tables import MatchTables


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(H.ROOT, "include", "specimux_b200.h")).read()
    declared = set(re.findall(r"\b(smx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), "libspecimux_b200.so does not export %s" % name
    assert set(_lib.EXPORTS) == declared
    assert lib.smx_abi_version() == 3


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.SmxParams) == 36
    assert _lib.RECORD_DTYPE.itemsize == 64 and _lib.RECORD32_DTYPE.itemsize == 32 and _lib.RECORD16_DTYPE.itemsize == 16
    assert ctypes.sizeof(_lib.SmxBatch) == 80 and ctypes.sizeof(_lib.SmxResults) == 112


@pytest.mark.skipif(_lib.load().smx_device_count() > 0, reason="box has a GPU")
def test_no_gpu_fails_loudly_no_cpu_fallback():
    g = H.load_golden("fixture")
    run = g["runs"]["default"]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    params = H.params_from_run(run, specimens)
    with pytest.raises(_lib.SmxError) as ei:
        Matcher(MatchTables(specimens, params))
    assert ei.value.code == _lib.SMX_ERR_NO_DEVICE
    with pytest.raises(_lib.SmxError):
        process_sequences(H.records(g["reads"]), params, specimens, H.make_args({}), H.prefilter_for(H.make_args({})))


def test_table_validation_errors():
    g = H.load_golden("fixture")
    run = g["runs"]["default"]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    params = H.params_from_run(run, specimens)
    params.max_dist_index = 13          # >= barcode length: rejected
    mt = MatchTables(specimens, params)
    lib = H.hostsim_binding()
    batch = PackedBatch(["ACGT" * 30])
    with pytest.raises(_lib.SmxError):
        Matcher(mt, binding=lib).match(batch)


@pytest.mark.parametrize("clip", [0, 80])
def test_packer_roundtrip(clip):
    rng = np.random.default_rng(5)
    reads = []
    for i in range(200):
        n = int(rng.integers(0, 400))
        s = "".join(rng.choice(list("ACGT"), size=n))
        if i % 7 == 0 and n:
            pos = rng.integers(0, n, size=3)
            s = "".join("NRUacx"[j % 6] if j in pos else c for j, c in enumerate(s))
        reads.append(s)
    b = PackedBatch(reads, clip=clip)
    assert b.n_reads == len(reads)
    code = {c: i for i, c in enumerate("ACGTRYSWKMBDHVN")}
    comp = reverse_complement
    for r, s in enumerate(reads):
        n = len(s)
        assert b.lengths[r] == n
        stored = list(range(n)) if not clip or n <= 2 * clip else list(range(clip)) + list(range(n - clip, n))
        flagged = b.n_flagged and b.off4[r] != np.uint64(2 ** 64 - 1)
        assert bool(flagged) == any(c not in "ACGT" for c in (s[i] for i in stored))
        if clip:        # clipped batches are packed at a fixed stride, 16-bit lengths beside the 32-bit ones
            assert b.word_off is None and b.stride == (2 * clip + 15) // 16 and b.lengths16[r] == n
        first_word = r * b.stride if b.stride else int(b.word_off[r])
        for si, i in enumerate(stored):
            w = b.packed2[first_word + si // 16]
            c2 = (int(w) >> (2 * (si % 16))) & 3
            assert c2 == (code[s[i]] if s[i] in "ACGT" else 0)
        if flagged:
            n4 = (len(stored) + 7) // 8
            rs = comp(s)
            for si, i in enumerate(stored):
                w = b.packed4[int(b.off4[r]) + si // 8]
                assert (int(w) >> (4 * (si % 8))) & 15 == code.get(s[i], 15)
                w = b.packed4[int(b.off4[r]) + n4 + si // 8]
                assert (int(w) >> (4 * (si % 8))) & 15 == code.get(rs[i], 15)


def test_output_tree_byte_identical_to_reference_expected_output(tmp_path):
    """Fixture through the kernel simulator + the product's writer == the reference's own
    tests/data/integration_test_suite/expected_output (15 FASTQ files)."""
    g = H.load_golden("fixture")
    run = g["runs"]["default"]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    args = H.make_args(run["flags"])
    params = H.params_from_run(run, specimens)
    ops, total, matched = process_sequences(H.records(g["reads"]), params, specimens, args, H.prefilter_for(args),
                                            None, 0, _binding=H.hostsim_binding())
    out = str(tmp_path / "out")
    with OutputManager(out, "", True) as om:
        for op in ops:
            om.write_sequence(op)
    expected = {k: v for k, v in g["expected_output"].items() if k.endswith(".fastq")}
    produced = {}
    for root, _d, files in os.walk(out):
        for f in files:
            p = os.path.join(root, f)
            produced[os.path.relpath(p, out)] = open(p).read()
    assert sorted(produced) == sorted(expected)
    for k in expected:
        assert produced[k] == expected[k], k


def test_parsers(tmp_path):
    p = tmp_path / "primers.fasta"
    p.write_text(">F1 pool=A,B position=forward\nacgtry\n>R1 pool=A;B position=reverse\nTTGACC\n")
    s = tmp_path / "specimens.txt"
    s.write_text("SampleID\tPrimerPool\tFwIndex\tFwPrimer\tRvIndex\tRvPrimer\nS1\tA\tacgtacgtacgta\tF1\tTTTTGGGGCCCCA\t*\n")
    reg = read_primers_file(str(p))
    assert reg.get_primer("F1").primer == "ACGTRY" and reg.get_primer("F1").primer_rc == "RYACGT"
    assert reg.get_primer("R1").pools == ["A", "B"]
    spec = read_specimen_file(str(s), reg)
    spec.validate()
    assert spec.b_length() == 13 and [x.name for x in spec.get_paired_primers("ACGTRY")] == ["R1"]
    assert reg.get_pools() == ["A"]          # pool B pruned: no specimen uses it
    fq = tmp_path / "r.fastq"
    fq.write_text("@id1 extra words\nACGT\n+\nIIII\n@id2\nAC\nGT\n+\nII\nII\n")
    recs = list(parse(str(fq), "fastq"))
    assert [(r.id, r.description, r.seq, r.qual) for r in recs] == [("id1", "id1 extra words", "ACGT", "IIII"),
                                                                    ("id2", "id2", "ACGT", "IIII")]


def test_simulator_seam_is_closed_outside_the_test_suite(monkeypatch):
    """The `binding` argument is a test seam: without SMX_TEST_SEAM=1 the product API refuses it (no CPU execution
    route through process_sequences / Matcher)."""
    from specimux_b200.engine import Matcher
    monkeypatch.delenv("SMX_TEST_SEAM", raising=False)
    with pytest.raises(RuntimeError, match="test seam"):
        Matcher(None, binding=object())


def test_file_pipeline_starts_one_gpu_per_share_of_input(tmp_path, monkeypatch):
    """orchestration._gpus_worth_starting: a further GPU per SMX_BYTES_PER_GPU bytes of input, never more than asked
    for, gzip input counted four-fold, unknown sizes left alone."""
    import argparse
    from specimux_b200 import orchestration
    f = tmp_path / "reads.fastq"
    f.write_bytes(b"x" * 3000)
    args = argparse.Namespace(sequence_file=str(f))
    monkeypatch.setenv("SMX_BYTES_PER_GPU", "1000")
    assert orchestration._gpus_worth_starting(args, 8) == 3
    assert orchestration._gpus_worth_starting(args, 2) == 2
    monkeypatch.setenv("SMX_BYTES_PER_GPU", str(8 << 30))
    assert orchestration._gpus_worth_starting(args, 8) == 1
    gz = tmp_path / "reads.fastq.gz"
    gz.write_bytes(b"x" * 1000)
    monkeypatch.setenv("SMX_BYTES_PER_GPU", "1000")
    assert orchestration._gpus_worth_starting(argparse.Namespace(sequence_file=str(gz)), 8) == 4
    assert orchestration._gpus_worth_starting(argparse.Namespace(sequence_file=str(tmp_path / "missing")), 5) == 5


def test_both_bench_arms_build_the_same_config_object():
    import bench
    a = bench.bench_config("ont037", 765_000, 80, 3)
    b = bench.bench_config("ont037", 765000, 80, 3)
    assert a == b and set(a) == {"workload", "description", "reads_per_gpu", "search_len", "k_index", "dereplicate", "trim", "l2"}
