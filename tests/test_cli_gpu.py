"""The reference's own black-box tests, run against this package's CLI on a GPU
(reference: tests/test_integration.py:75-140, tests/test_orientation_normalization.py:27-204),
strengthened to a byte-for-byte comparison with the reference's expected_output tree."""
import os
import re
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _write_inputs(tmp, golden):
    g = H.load_golden(golden)
    p, s, q = tmp / "primers.fasta", tmp / "specimens.txt", tmp / "sequences.fastq"
    with open(p, "w") as fh:
        for name, seq, pos, pools in g["primers"]:
            fh.write(">%s pool=%s position=%s\n%s\n" % (name, ",".join(pools), pos, seq))
    with open(s, "w") as fh:
        fh.write("SampleID\tPrimerPool\tFwIndex\tFwPrimer\tRvIndex\tRvPrimer\n")
        for row in g["specimens"]:
            fh.write("\t".join(row) + "\n")
    with open(q, "w") as fh:
        for rid, seq, qual in g["reads"]:
            fh.write("@%s\tdx:i:0\n%s\n+\n%s\n" % (rid, seq, qual))
    return g, str(p), str(s), str(q)


def _run_cli(*argv):
    return subprocess.run([sys.executable, "-m", "specimux_b200.cli", *argv], capture_output=True, text=True,
                          cwd=H.ROOT, timeout=600)


def _tree(out):
    files = {}
    for root, _d, fs in os.walk(out):
        for f in fs:
            p = os.path.join(root, f)
            files[os.path.relpath(p, out)] = open(p).read()
    return files


def test_full_pipeline_matches_reference_expected_output(tmp_path):
    g, p, s, q = _write_inputs(tmp_path, "fixture")
    out = str(tmp_path / "out")
    r = _run_cli(p, s, q, "-F", "-O", out, "-d")
    assert r.returncode == 0, r.stderr
    assert "Processed 40 sequences" in r.stderr
    m = re.search(r"match rate: ([\d.]+)%", r.stderr)
    assert m and abs(float(m.group(1)) - 15.0) <= 5.0
    for d in ("full", "partial", "unknown", "trace"):
        assert os.path.isdir(os.path.join(out, d))
    assert [f for f in os.listdir(os.path.join(out, "trace")) if f.endswith(".tsv")]
    produced = {k: v for k, v in _tree(out).items() if not k.startswith("trace") and k != "log.txt"}
    expected = g["expected_output"]
    assert sorted(produced) == sorted(expected)
    for k in expected:
        assert produced[k] == expected[k], k


@pytest.mark.parametrize("native", ["1", "0"])
def test_full_pipeline_both_io_routes_match_reference_expected_output(tmp_path, native):
    """No -d: the native reader/writer route (SMX_NATIVE_IO=1, default) and the Python-object route
    (SMX_NATIVE_IO=0) both reproduce the reference's expected_output tree byte for byte."""
    g, p, s, q = _write_inputs(tmp_path, "fixture")
    out = str(tmp_path / "out")
    env = dict(os.environ, SMX_NATIVE_IO=native)
    r = subprocess.run([sys.executable, "-m", "specimux_b200.cli", p, s, q, "-F", "-O", out], capture_output=True,
                       text=True, cwd=H.ROOT, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert "Processed 40 sequences" in r.stderr
    assert ("native reader/writer" in r.stderr) == (native == "1")
    produced = {k: v for k, v in _tree(out).items() if k != "log.txt"}
    expected = g["expected_output"]
    assert sorted(produced) == sorted(expected)
    for k in expected:
        assert produced[k] == expected[k], k


def test_native_route_gzip_input_and_console_output(tmp_path):
    import gzip
    g, p, s, q = _write_inputs(tmp_path, "fixture")
    with open(q, "rb") as src, gzip.open(q + ".gz", "wb") as dst:
        dst.write(src.read())
    a = _run_cli(p, s, q, "-n", "12")
    b = _run_cli(p, s, q + ".gz", "-n", "12")
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout == b.stdout and a.stdout.count("\n") >= 48
    env = dict(os.environ, SMX_NATIVE_IO="0")
    c = subprocess.run([sys.executable, "-m", "specimux_b200.cli", p, s, q, "-n", "12"], capture_output=True, text=True,
                       cwd=H.ROOT, timeout=600, env=env)
    assert c.returncode == 0 and c.stdout == a.stdout


@pytest.mark.parametrize("n", [5, 10, 20])
def test_partial_sequences(tmp_path, n):
    g, p, s, q = _write_inputs(tmp_path, "fixture")
    r = _run_cli(p, s, q, "-F", "-O", str(tmp_path / "out"), "-n", str(n))
    assert r.returncode == 0, r.stderr
    assert "Processed %d sequences" % n in r.stderr
    m = re.search(r"match rate: ([\d.]+)%", r.stderr)
    assert m and abs(float(m.group(1)) - 20.0) <= 5.0


def test_orientation_normalization(tmp_path):
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    _g, p, s, q = _write_inputs(tmp_path / "a", "fixture")
    _g2, p2, s2, q2 = _write_inputs(tmp_path / "b", "fixture_rc")
    o1, o2 = str(tmp_path / "o1"), str(tmp_path / "o2")
    assert _run_cli(p, s, q, "-F", "-O", o1).returncode == 0
    assert _run_cli(p2, s2, q2, "-F", "-O", o2).returncode == 0
    t1 = {k: v for k, v in _tree(o1).items() if k.startswith("full") and k.endswith(".fastq")}
    t2 = {k: v for k, v in _tree(o2).items() if k.startswith("full") and k.endswith(".fastq")}
    common = set(t1) & set(t2)
    assert common
    for k in common:      # first record of each per-specimen file carries the same (normalised) bases
        assert t1[k].split("\n")[1] == t2[k].split("\n")[1], k


def test_console_output_and_version(tmp_path):
    _g, p, s, q = _write_inputs(tmp_path, "fixture")
    r = _run_cli(p, s, q, "-n", "5")
    assert r.returncode == 0 and r.stdout.count("\n") == 20
    v = _run_cli("--version")
    assert v.returncode == 0 and "specimux version" in v.stdout


def test_two_gpus_give_the_same_tree_as_one(tmp_path):
    """`-t 2` deals batches round-robin to two GPUs (one feeder thread each) and writes in submission order:
    the output tree must be byte-identical to the one-GPU run, on both I/O routes."""
    from specimux_b200 import _lib, synth
    if _lib.load().smx_device_count() < 2:
        pytest.skip("needs two GPUs")
    ds = synth.ont037(n_reads=30_000, seed=11)
    p, s, q = str(tmp_path / "primers.fasta"), str(tmp_path / "specimens.txt"), str(tmp_path / "reads.fastq")
    ds.write_tables(p, s)
    ds.write_fastq(q)
    trees = {}
    for route in ("1", "0"):
        for t in ("1", "2"):
            out = str(tmp_path / ("out_%s_%s" % (route, t)))
            # SMX_BYTES_PER_GPU=1: start every GPU asked for (the default policy starts one per 8 GiB of input)
            env = dict(os.environ, SMX_NATIVE_IO=route, SMX_GPU_BATCH_READS="4096", SMX_BYTES_PER_GPU="1",
                       SMX_READER_CHUNK_BYTES=str(1 << 20))
            r = subprocess.run([sys.executable, "-m", "specimux_b200.cli", p, s, q, "-F", "-O", out, "-t", t],
                               capture_output=True, text=True, cwd=H.ROOT, timeout=900, env=env)
            assert r.returncode == 0, r.stderr
            assert ("Will run on %s GPU(s)" % t) in r.stderr
            trees[(route, t)] = {k: v for k, v in _tree(out).items() if k != "log.txt"}
    ref = trees[("1", "1")]
    assert len(ref) > 700
    for key, tree in trees.items():
        assert tree == ref, key
