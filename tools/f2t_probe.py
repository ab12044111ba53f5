#!/usr/bin/env python3
"""File -> tree probe: the bench's ont037 FASTQ through the CLI with -t 1 .. -t N, SMX_IO_TRACE=1 (busy seconds per
pipeline stage), each twice; prints the CLI's log lines.  usage: python tools/f2t_probe.py [n_gpus ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (bench._workload_files puts the files on tmpfs when there is room)

# arguments: GPU counts, and VAR=value settings tried one after the other on top of the environment ("-" = none)
gpus = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1]
settings = [a for a in sys.argv[1:] if not a.isdigit()] or ["-"]
n = 765_000
ds = bench.dataset("ont037", n, 0)
d, files = bench._workload_files("ont037", ds, n)
for g, setting in [(g, st) for g in gpus for st in settings]:
    extra = dict(kv.split("=", 1) for kv in setting.split(",") if "=" in kv)
    print("#### -t %d %s" % (g, setting), flush=True)
    for rep in range(2):
        out = os.path.join(d, "out_probe")
        subprocess.run(["rm", "-rf", out])
        env = dict(os.environ, PYTHONPATH=ROOT, SMX_IO_TRACE="1", HOME=d, **extra)
        t0 = time.perf_counter()
        r = subprocess.run([sys.executable, "-m", "specimux.cli"] + files + ["-F", "-O", out, "-t", str(g)] ,
                           capture_output=True, text=True, env=env, cwd=d)
        print("== -t %d run %d: rc %d wall %.2fs" % (g, rep, r.returncode, time.perf_counter() - t0), flush=True)
        for line in r.stderr.splitlines():
            if "INFO" in line or "WARN" in line or "rror" in line:
                print("   ", line[24:220], flush=True)
