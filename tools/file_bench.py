#!/usr/bin/env python3
"""File-to-tree throughput: synthetic FASTQ on local disk -> finished per-specimen output tree through the CLI
(SURVEY.md 8d "end-to-end reads/s from FASTQ on local disk to finished output tree").  Runs the native
reader/writer route and, on a smaller slice, the Python-object route of the same package for comparison.
usage (GPU box): python tools/file_bench.py [--config ont037] [--reads 765000] [--py-reads 50000] [--gz]"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def tree_bytes(d):
    n = b = 0
    for root, _s, fs in os.walk(d):
        for f in fs:
            if f.endswith((".fastq", ".fasta")):
                n += 1
                b += os.path.getsize(os.path.join(root, f))
    return n, b


def run_cli(p, s, q, out, env_extra, extra=()):
    env = dict(os.environ, **env_extra)
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "specimux_b200.cli", p, s, q, "-F", "-O", out, *extra],
                       cwd=ROOT, env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit("cli failed: " + r.stderr[-2000:])
    inner = [l for l in r.stderr.splitlines() if "Elapsed time" in l or "Processed" in l]
    return dt, inner


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="ont037")
    ap.add_argument("--reads", type=int, default=765000)
    ap.add_argument("--py-reads", type=int, default=50000)
    ap.add_argument("--gz", action="store_true")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    from specimux_b200 import synth
    tmp = tempfile.mkdtemp(prefix="smx_file_bench_")
    t0 = time.time()
    ds = synth.CONFIGS[a.config](n_reads=a.reads)
    p, s = os.path.join(tmp, "primers.fasta"), os.path.join(tmp, "specimens.txt")
    q = os.path.join(tmp, "reads.fastq")
    ds.write_tables(p, s)
    ds.write_fastq(q)
    if a.py_reads:
        q_small = os.path.join(tmp, "reads_small.fastq")
        ds.write_fastq(q_small, 0, a.py_reads)
    size = os.path.getsize(q)
    if a.gz:
        subprocess.check_call(["gzip", "-1", "-k", q])
        q += ".gz"
    print("generated %d reads (%.0f MB FASTQ) in %.1fs" % (a.reads, size / 1e6, time.time() - t0), file=sys.stderr)
    out = {"config": a.config, "reads": a.reads, "fastq_mb": size / 1e6, "gz": a.gz}
    # native route, twice (first run pays CUDA context + page-cache warm-up)
    for rep in range(2):
        o = os.path.join(tmp, "out_native_%d" % rep)
        dt, inner = run_cli(p, s, q, o, {"SMX_NATIVE_IO": "1"})
        nf, nb = tree_bytes(o)
        out["native_run%d" % rep] = {"wall_s": dt, "reads_per_s_wall": a.reads / dt, "log": inner, "files": nf,
                                     "tree_mb": nb / 1e6}
        if not a.keep:
            shutil.rmtree(o)
    if a.py_reads:
        o = os.path.join(tmp, "out_py")
        dt, inner = run_cli(p, s, q_small, o, {"SMX_NATIVE_IO": "0"})
        out["python_route"] = {"reads": a.py_reads, "wall_s": dt, "reads_per_s_wall": a.py_reads / dt, "log": inner}
        o2 = os.path.join(tmp, "out_native_small")
        dt2, inner2 = run_cli(p, s, q_small, o2, {"SMX_NATIVE_IO": "1"})
        out["native_small"] = {"reads": a.py_reads, "wall_s": dt2, "log": inner2}
        # the two routes must produce the same tree
        same = subprocess.run(["diff", "-r", "-q", "-x", "log.txt", o, o2], capture_output=True, text=True)
        out["routes_identical"] = same.returncode == 0
        if same.returncode != 0:
            out["diff"] = same.stdout[:2000]
    print(json.dumps(out, indent=1))
    if not a.keep:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
