"""Native reader / writer (libspecimux_io.so) against the Python implementations
of the same reference behaviour (seqio.py = Bio.SeqIO as specimux uses it; io_utils.OutputManager =
reference io_utils.py:179-268) and against the reference's own expected_output tree.  The matching
between them runs on the CPU kernel simulator here; tests/test_cli_gpu.py runs the same route on a GPU."""
import gzip
import os

import numpy as np
import pytest

import helpers as H
from specimux_b200 import native_io, seqio
from specimux_b200.demultiplex import get_matcher, process_sequences
from specimux_b200.engine import PackedBatch
from specimux_b200.io_utils import OutputManager, output_write_operation

FASTQ_CASES = {
    "plain": "@r1 desc one\nACGTACGT\n+\nIIIIHHHH\n@r2\tx:i:1\nTTTT\n+r2\n!!!!\n",
    "no_trailing_newline": "@r1\nACGT\n+\nIIII",
    "crlf": "@r1 a b\r\nACGTNN\r\n+\r\nIIIIII\r\n@r2\r\nAC\r\n+\r\nII\r\n",
    "blank_lines": "\n\n@r1\nACGT\n+\nIIII\n\n   \n@r2\nGG\n+\nII\n\n",
    "multi_line": "@r1 wrapped\nACGT\nACGT\nAC\n+\nIIII\nIIII\nII\n@r2\nTT\n+\nII\n",
    "quality_starts_with_at": "@r1\nACGT\n+\n@III\n@r2\nACGT\n+\n+III\n",
    "lower_and_iupac": "@r1\nacgtNRYKMswbdhvU\n+\nIIIIIIIIIIIIIIII\n",
    "title_whitespace": "@  lead  mid \t tail  \nAC\n+\nII\n@\nGT\n+\nII\n@x\xc2\xa0y z\nAA\n+\nII\n",
    "empty_sequence": "@r1\n\n+\n\n@r2\nAC\n+\nII\n",
    "padded_lines": "@r1\n  ACGT  \n+\n  IIII\t\n",
}

FASTA_CASES = {
    "plain": ">s1 first\nACGT\nACGT\n>s2\nTT\n",
    "leading_junk": "junk line\n\n>s1\nAC\n\nGT\n>s2 x\n",
    "crlf": ">s1 d\r\nACGT\r\nAC\r\n>s2\r\nGG\r\n",
    "no_trailing_newline": ">s1\nACGT",
    "spaces": ">  s1  pad  \n  AC GT  \nTT\n",
}


def _native_records(path, is_fastq, per_block=3):
    out = []
    with native_io.FastxReader(path, is_fastq) as rd:
        blk = native_io.ReadBlock()
        while True:
            rd.next_block(per_block, blk)
            if blk.n_reads == 0:
                break
            out.extend(blk.read(r) for r in range(blk.n_reads))
    return out


def _python_records(path, fmt):
    return [(r.id, r.description, r.seq, r.qual) for r in seqio.parse(path, fmt)]


def test_library_exports_every_declared_symbol():
    import re
    lib = native_io.load()
    header = open(os.path.join(H.ROOT, "include", "specimux_io.h")).read()
    declared = set(re.findall(r"\b(smx_[a-z0-9_]+)\s*\(", header))
    assert declared and set(native_io.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), "libspecimux_io.so does not export %s" % name
    assert lib.smx_io_abi_version() == 2


@pytest.mark.parametrize("name", sorted(FASTQ_CASES))
@pytest.mark.parametrize("gz", [False, True])
def test_fastq_reader_matches_python_parser(tmp_path, name, gz):
    path = str(tmp_path / ("x.fastq.gz" if gz else "x.fastq"))
    data = FASTQ_CASES[name].encode("latin-1")
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as fh:
        fh.write(data)
    assert _native_records(path, True) == _python_records(path, "fastq")


@pytest.mark.parametrize("name", sorted(FASTA_CASES))
def test_fasta_reader_matches_python_parser(tmp_path, name):
    path = str(tmp_path / "x.fasta")
    with open(path, "wb") as fh:
        fh.write(FASTA_CASES[name].encode("latin-1"))
    got = _native_records(path, False)
    assert got == _python_records(path, "fasta")
    assert all(q is None for *_x, q in got)


@pytest.mark.parametrize("text,message", [
    ("ACGT\n+\nIIII\n", "Records in Fastq files should start with '@' character"),
    ("@r1\nACGT\n+\nIII\n", "Lengths of sequence and quality values differs"),
    ("@r1\nACGT\n+\nIIII\n@r2\nACGT\n+\nII", "Lengths of sequence and quality values differs"),
])
def test_fastq_reader_errors_like_the_python_parser(tmp_path, text, message):
    path = str(tmp_path / "bad.fastq")
    open(path, "w").write(text)
    with pytest.raises(ValueError, match=message):
        _python_records(path, "fastq")
    with pytest.raises(ValueError, match=message):
        _native_records(path, True)


def test_reader_missing_file():
    with pytest.raises(OSError):
        native_io.FastxReader("/nonexistent/x.fastq", True)


def test_reader_large_random_file_block_boundaries(tmp_path):
    """Records straddling the 8 MB refill buffer, long lines, skip()."""
    rng = np.random.default_rng(11)
    path = str(tmp_path / "big.fastq")
    expect = []
    with open(path, "w") as fh:
        for i in range(3000):
            n = int(rng.integers(0, 9000))
            seq = "".join(rng.choice(list("ACGTN"), size=n))
            qual = "".join(chr(33 + int(q)) for q in rng.integers(0, 41, size=n))
            title = "read%d runid=%d" % (i, rng.integers(0, 10 ** 9))
            fh.write("@%s\n%s\n+\n%s\n" % (title, seq, qual))
            expect.append(("read%d" % i, title, seq, qual))
    assert os.path.getsize(path) > 24 << 20
    assert _native_records(path, True, per_block=700) == expect
    with native_io.FastxReader(path, True) as rd:
        assert rd.skip(2990) == 2990
        blk = rd.next_block(100)
        assert [blk.read(r) for r in range(blk.n_reads)] == expect[2990:]
        assert rd.skip(5) == 0


def _tree(root):
    files = {}
    for d, _s, fs in os.walk(root):
        for f in fs:
            p = os.path.join(d, f)
            files[os.path.relpath(p, root)] = open(p, "rb").read()
    return files


def _write_fastq(path, reads, fasta=False):
    with open(path, "w") as fh:
        for rid, seq, qual in reads:
            if fasta:
                fh.write(">%s extra words\n%s\n" % (rid, seq))
            else:
                fh.write("@%s extra words\n%s\n+\n%s\n" % (rid, seq, qual))


def _golden_runs():
    for name in H.golden_names():
        g = H.load_golden(name)
        for run in g["runs"]:
            yield name, run


@pytest.mark.parametrize("name,run_name", list(_golden_runs()))
def test_native_writer_tree_equals_python_output_manager(tmp_path, name, run_name):
    """Same records through both writers: reader -> packer -> (simulated) matching -> native writer versus
    process_sequences -> WriteOperation -> OutputManager.  Every file identical, every golden run
    (all trim modes, dereplicate best/none, partial / unknown / multiple-specimen records)."""
    g = H.load_golden(name)
    run = g["runs"][run_name]
    reads = [tuple(r) for r in g["reads"]][:400]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    args = H.make_args(run["flags"])
    args.output_dir, args.output_file_prefix = str(tmp_path / "py"), "pre_"
    params = H.params_from_run(run, specimens)
    binding = H.hostsim_binding()
    # Python route
    fq = str(tmp_path / "in.fastq")
    _write_fastq(fq, reads)
    ops, _n, _m = process_sequences(list(seqio.parse(fq, "fastq")), params, specimens, args, H.prefilter_for(args),
                                    None, 0, _binding=binding)
    with OutputManager(args.output_dir, args.output_file_prefix, True) as om:
        for op in ops:
            output_write_operation(op, om, args)
    # native route
    matcher = get_matcher(params, specimens, args, H.prefilter_for(args), 0, binding)
    native_dir = str(tmp_path / "native")
    with native_io.FastxReader(fq, True) as rd, native_io.TreeWriter(native_dir, "pre_", True, matcher.tables) as wr:
        blk = native_io.ReadBlock()
        n_rec = 0
        while rd.next_block(97, blk).n_reads:
            res = matcher.match(PackedBatch.from_block(blk, clip=params.search_len))
            wr.write(blk, res.records)
            n_rec += len(res.records)
        assert wr.stats()[0] == n_rec == len(ops)
    assert _tree(native_dir) == _tree(args.output_dir)
    # the compact record form (smx_record32, no location pairs) gives the same tree
    compact_dir = str(tmp_path / "compact")
    with native_io.FastxReader(fq, True) as rd, native_io.TreeWriter(compact_dir, "pre_", True, matcher.tables) as wr:
        blk = native_io.ReadBlock()
        while rd.next_block(97, blk).n_reads:
            res = matcher.match(PackedBatch.from_block(blk, clip=params.search_len), compact=True)
            assert res.records.dtype.itemsize == 32
            wr.write(blk, res.records)
    assert _tree(compact_dir) == _tree(args.output_dir)
    # ... and so does the 16-byte wire form (smx_record16: no read index, no offsets, 16-bit trim extents)
    wire_dir = str(tmp_path / "wire")
    with native_io.FastxReader(fq, True) as rd, native_io.TreeWriter(wire_dir, "pre_", True, matcher.tables) as wr:
        blk = native_io.ReadBlock()
        while rd.next_block(97, blk).n_reads:
            res = matcher.match(PackedBatch.from_block(blk, clip=params.search_len), compact="wire")
            assert res.records.dtype.itemsize == 16 and res.rec_offset is None
            wr.write(blk, res.records)
    assert _tree(wire_dir) == _tree(args.output_dir)


def test_native_writer_fasta_input_and_console_form(tmp_path, capfd):
    g = H.load_golden("fixture")
    run = g["runs"]["default"]
    reads = [tuple(r) for r in g["reads"]]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    args = H.make_args(run["flags"])
    params = H.params_from_run(run, specimens)
    binding = H.hostsim_binding()
    fa = str(tmp_path / "in.fasta")
    _write_fastq(fa, reads, fasta=True)
    matcher = get_matcher(params, specimens, args, H.prefilter_for(args), 0, binding)
    ops, _n, _m = process_sequences(list(seqio.parse(fa, "fasta")), params, specimens, args, H.prefilter_for(args),
                                    None, 0, _binding=binding)
    # FASTA tree
    args.output_dir, args.isfastq = str(tmp_path / "py"), False
    with OutputManager(args.output_dir, "", False) as om:
        for op in ops:
            output_write_operation(op, om, args)
    with native_io.FastxReader(fa, False) as rd, native_io.TreeWriter(str(tmp_path / "nat"), "", False, matcher.tables) as wr:
        blk = rd.next_block(1000)
        res = matcher.match(PackedBatch.from_block(blk, clip=params.search_len))
        wr.write(blk, res.records)
    assert _tree(str(tmp_path / "nat")) == _tree(args.output_dir)
    # console form, FASTQ symbol with synthesised Q40 qualities (alignment.py:52-56)
    capfd.readouterr()
    args.output_to_files, args.isfastq = False, True
    import sys
    for op in ops:
        output_write_operation(op, None, args)
    sys.stdout.flush()
    expected = capfd.readouterr().out
    with native_io.FastxReader(fa, False) as rd, native_io.TreeWriter(None, "", True, matcher.tables) as wr:
        blk = rd.next_block(1000)
        res = matcher.match(PackedBatch.from_block(blk, clip=params.search_len))
        wr.write(blk, res.records)
    assert capfd.readouterr().out == expected
    assert expected.count("\n") == 4 * len(ops)


@pytest.mark.parametrize("gz", [False, True])
def test_native_pipeline_reproduces_reference_expected_output(tmp_path, gz):
    """File -> tree through orchestration._run_native (threads: reader+packer, matcher, writer) equals the
    reference's own tests/data/integration_test_suite/expected_output, byte for byte."""
    import argparse
    from specimux_b200 import orchestration
    g = H.load_golden("fixture")
    run = g["runs"]["default"]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    params = H.params_from_run(run, specimens)
    fq = str(tmp_path / ("sequences.fastq.gz" if gz else "sequences.fastq"))
    with (gzip.open(fq, "wt") if gz else open(fq, "w")) as fh:
        for rid, seq, qual in g["reads"]:
            fh.write("@%s\tdx:i:0\n%s\n+\n%s\n" % (rid, seq, qual))
    out = str(tmp_path / "out")
    args = argparse.Namespace(**vars(H.make_args(run["flags"])))
    args.sequence_file, args.output_dir, args.output_file_prefix = fq, out, ""
    args.num_seqs, args.start_seq, args.output_to_files = -1, 1, True
    orchestration.create_output_files(args, specimens)
    old = orchestration.GPU_BATCH_READS
    orchestration.GPU_BATCH_READS = 7            # many small batches through the job ring
    try:
        total, matched = orchestration._run_native(args, specimens, params, 1, H.prefilter_for(args),
                                                   _binding=H.hostsim_binding())
    finally:
        orchestration.GPU_BATCH_READS = old
    orchestration.cleanup_empty_directories(out)
    assert (total, matched) == (run["total"], run["matched"])
    produced = {k: v.decode() for k, v in _tree(out).items()}
    expected = g["expected_output"]
    assert sorted(produced) == sorted(expected)
    for k in expected:
        assert produced[k] == expected[k], k


@pytest.mark.parametrize("chunk,threads", [(1 << 16, 3), (70_001, 5), (1 << 18, 2)])
def test_parallel_reader_gives_the_serial_tree(tmp_path, monkeypatch, chunk, threads):
    """K parser threads over byte ranges of the FASTQ file (re-synchronised on record boundaries, quality lines that
    begin with '@' included) produce the same tree, in the same per-file record order, as the serial reader."""
    import argparse
    from specimux_b200 import orchestration, synth
    ds = synth.ont037(n_reads=700, seed=11)
    fq = str(tmp_path / "reads.fastq")
    with open(fq, "w") as fh:
        for i, (rid, seq, qual) in enumerate(ds.reads()):
            if i % 3 == 0:
                qual = "@" + qual[1:]            # a quality line starting with '@' must not be taken for a title
            if i % 5 == 0:
                qual = "+" + qual[1:]
            fh.write("@%s extra=%d\n%s\n+\n%s\n" % (rid, i, seq, qual))
    specimens = H.build_specimens(ds.primers, ds.specimens)
    from oracle import pipeline as orc
    oparams = orc.setup_params(orc.Tables(ds.primers, ds.specimens))
    from specimux_b200.models import MatchParameters
    params = MatchParameters(dict(oparams.max_dist_primers), oparams.max_dist_index, 80, True)
    trees = {}
    for mode in ("serial", "parallel"):
        args = argparse.Namespace(**vars(H.make_args({})))
        args.sequence_file, args.output_dir, args.output_file_prefix = fq, str(tmp_path / mode), ""
        args.num_seqs, args.start_seq, args.output_to_files = -1, 1, True
        monkeypatch.setenv("SMX_PARALLEL_READER", "0" if mode == "serial" else "1")
        monkeypatch.setenv("SMX_READER_CHUNK_BYTES", str(chunk))
        monkeypatch.setenv("SMX_READER_THREADS", str(threads))
        assert bool(orchestration._parallel_read_ok(args, fq, True)) == (mode == "parallel")
        total, matched = orchestration._run_native(args, specimens, params, 1, H.prefilter_for(args),
                                                   _binding=H.hostsim_binding())
        assert total == 700
        trees[mode] = (_tree(args.output_dir), matched)
    assert trees["serial"][1] == trees["parallel"][1]
    assert trees["serial"][0] == trees["parallel"][0]
    assert len(trees["serial"][0]) > 50


def test_range_readers_partition_the_file(tmp_path):
    """Consecutive byte ranges see every record exactly once, whatever the cut points."""
    rng = np.random.default_rng(3)
    recs = []
    for i in range(400):
        n = int(rng.integers(1, 300))
        seq = "".join(rng.choice(list("ACGT"), size=n))
        qual = "".join(rng.choice(list("@+!IJK#5"), size=n))
        recs.append(("r%d" % i, seq, qual))
    fq = str(tmp_path / "x.fastq")
    _write_fastq(fq, recs)
    size = os.path.getsize(fq)
    for step in (97, 1000, 4096, size // 3 + 1, size + 10):
        got = []
        for lo in range(0, size, step):
            with native_io.FastxReader(fq, True, byte_range=(lo, lo + step)) as rd:
                blk = rd.next_block(1 << 30)
                got.extend(blk.read(i)[0] for i in range(blk.n_reads))
        assert got == [r[0] for r in recs], step


def test_native_pipeline_num_seqs_and_start(tmp_path):
    import argparse
    from specimux_b200 import orchestration
    g = H.load_golden("fixture")
    run = g["runs"]["default"]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    params = H.params_from_run(run, specimens)
    fq = str(tmp_path / "sequences.fastq")
    _write_fastq(fq, [tuple(r) for r in g["reads"]])
    args = argparse.Namespace(**vars(H.make_args(run["flags"])))
    args.sequence_file, args.output_dir, args.output_file_prefix = fq, str(tmp_path / "o"), ""
    args.num_seqs, args.start_seq, args.output_to_files = 10, 6, True
    total, _m = orchestration._run_native(args, specimens, params, 1, H.prefilter_for(args), _binding=H.hostsim_binding())
    assert total == 10
    ids = set()
    for v in _tree(args.output_dir).values():
        ids.update(l.split()[0][1:] for l in v.decode().split("\n")[0::4] if l)
    assert ids <= {r[0] for r in g["reads"][5:15]} and ids


def test_fastq_reader_agrees_with_python_parser_on_random_text(tmp_path):
    """Random files out of well-formed and damaged FASTQ records: both parsers return the same records or both
    raise ValueError (the native one never crashes, never silently differs)."""
    rng = np.random.default_rng(123)

    def rnd(n, alphabet):
        return "".join(rng.choice(list(alphabet), size=n)) if n else ""

    def record():
        n = int(rng.integers(0, 40))
        seq, qual = rnd(n, "ACGTNacgu"), rnd(n, "!#5?I~")
        # ASCII only here: the native reader compares sequence / quality lengths in BYTES, Python in characters, which
        # differs only when a damaged file makes a non-ASCII title line part of a sequence or quality (documented)
        title = rng.choice(["r%d" % rng.integers(0, 99), "r x\ty", "", " lead", "a b  c "])
        form = rng.random()
        if form < 0.55:
            lines = ["@" + title, seq, "+", qual]
        elif form < 0.65 and n > 4:                              # wrapped sequence and quality
            k = int(rng.integers(1, n))
            lines = ["@" + title, seq[:k], seq[k:], "+" + title, qual[:k], qual[k:]]
        elif form < 0.72:
            lines = ["@" + title, seq + " ", "+", " " + qual, "", "  "]      # padding, blank lines after
        elif form < 0.80:
            lines = ["@" + title, seq, "+", qual[:-1] if n else "I"]          # length mismatch
        elif form < 0.86:
            lines = [title, seq, "+", qual]                                 # no '@'
        elif form < 0.92:
            lines = ["@" + title, seq, qual]                                # no '+' line
        else:
            lines = ["@" + title, seq, "+", "@" + qual[1:] if n else ""]      # quality starting with '@'
        return lines

    agree = raised = 0
    for trial in range(300):
        lines = []
        for _ in range(int(rng.integers(1, 6))):
            lines += record()
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        text = eol.join(lines) + (eol if rng.random() < 0.7 else "")
        path = str(tmp_path / ("f%d.fastq" % trial))
        with open(path, "wb") as fh:
            fh.write(text.encode("latin-1"))
        try:
            want = _python_records(path, "fastq")
        except ValueError:
            with pytest.raises(ValueError):
                _native_records(path, True)
            raised += 1
            continue
        assert _native_records(path, True) == want, repr(text)
        agree += 1
    assert agree > 60 and raised > 60, (agree, raised)


def test_compact_records_are_the_projection_of_full_records():
    """smx_record32 == smx_record minus the four location pairs, field by field (kernel simulator here; the CUDA
    library's k_pack_records32 is compared against the same simulator output in the GPU tier through the CLI trees)."""
    g = H.load_golden("synth_dense")
    run = g["runs"]["derep_none"]
    specimens = H.build_specimens(g["primers"], g["specimens"])
    params = H.params_from_run(run, specimens)
    args = H.make_args(run["flags"])
    matcher = get_matcher(params, specimens, args, H.prefilter_for(args), 0, H.hostsim_binding())
    batch = PackedBatch([r[1] for r in g["reads"]], clip=params.search_len)
    full = matcher.match(batch).records
    lite = matcher.match(batch, compact=True).records
    assert len(full) == len(lite) > len(g["reads"]) // 2
    for name in ("read", "sample", "trim_start", "trim_end", "pool", "p1", "p2", "resolution", "candidate"):
        assert np.array_equal(full[name], lite[name]), name
    assert np.array_equal(full["dist"], lite["dist"])
    assert np.array_equal(lite["flags"] & 1, full["reverse"]) and np.array_equal((lite["flags"] >> 1) & 1, full["trim_empty"])
    # the 16-byte wire form: read index from the last-of-read flags, trim_end from the read length
    wire = matcher.match(batch, compact="wire").records
    assert len(wire) == len(full)
    last = (wire["flags"] >> 5) & 1
    read = np.concatenate([[0], np.cumsum(last)[:-1]]).astype(np.int64)
    assert np.array_equal(read, full["read"])
    lens = np.array([len(r[1]) for r in g["reads"]])[read]
    assert np.array_equal(wire["sample"], full["sample"]) and np.array_equal(wire["trim_start"], full["trim_start"])
    assert np.array_equal(lens - wire["trim_tail"], full["trim_end"]) and np.array_equal(wire["pool"], full["pool"])
    for a, b in (("p1", "p1"), ("p2", "p2")):
        assert np.array_equal(np.where(wire[a] == 255, -1, wire[a].astype(int)), full[b])
    d = full["dist"].astype(int)
    assert np.array_equal(np.where(wire["dist_p1"] == 255, -1, wire["dist_p1"].astype(int)), d[:, 0])
    assert np.array_equal(np.where(wire["dist_p2"] == 255, -1, wire["dist_p2"].astype(int)), d[:, 3])
    b1, b2 = wire["dist_b"] & 15, wire["dist_b"] >> 4
    assert np.array_equal(np.where(b1 == 15, -1, b1.astype(int)), d[:, 1]) and np.array_equal(np.where(b2 == 15, -1, b2.astype(int)), d[:, 2])
    assert np.array_equal(wire["flags"] & 7, full["resolution"]) and np.array_equal((wire["flags"] >> 3) & 1, full["reverse"])
    assert np.array_equal((wire["flags"] >> 4) & 1, full["trim_empty"])
