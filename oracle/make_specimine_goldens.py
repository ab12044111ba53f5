#!/usr/bin/env python3
"""oracle/make_specimine_goldens.py -- TEST INFRASTRUCTURE.  Run in the BUILD container only (needs /root/reference).

Builds small specimux output trees (full / partial FASTQ files of one specimen), runs the UNMODIFIED reference
`specimine` (src/specimux/specimine.py through specimux.cli:specimine_main) over the stand-ins of oracle/standins
(edlib -> oracle/edlib_restated.c, Bio) and stores inputs + the `.mined` file it wrote in
tests/golden_specimine/cases.json.gz.  The product (specimux_b200/specimine.py) must write the same bytes."""
import gzip
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
OUT = os.path.join(ROOT, "tests", "golden_specimine", "cases.json.gz")


def mutate(rng, s, rate):
    out = []
    for ch in s:
        u = rng.random()
        if u < rate * 0.4:
            out.append(rng.choice([c for c in "ACGT" if c != ch]))
        elif u < rate * 0.7:
            out.append(ch)
            out.append(rng.choice("ACGT"))
        elif u < rate:
            continue
        else:
            out.append(ch)
    return "".join(out)


def rnd(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def fastq(records):
    return "".join("@%s\n%s\n+\n%s\n" % (t, s, q) for t, s, q in records)


def make_case(seed, amplicon_len, n_full, n_partial, level):
    rng = random.Random(seed)
    amplicon = rnd(rng, amplicon_len)
    other = rnd(rng, amplicon_len)
    b1, b2 = rnd(rng, 13), rnd(rng, 13)
    qual = lambda s: "".join(chr(33 + rng.randint(5, 40)) for _ in s)
    full = []
    for i in range(n_full):
        s = mutate(rng, amplicon, rng.choice([0.0, 0.02, 0.05, 0.08]))
        full.append(("full%02d 0,1,0,2 pool=ITS primers=ITS1F+ITS4 SP1" % i, s, qual(s)))
    partials = {}
    for kind in ("fwd", "rev"):
        recs = []
        for i in range(n_partial):
            r = rng.random()
            if r < 0.45:       # the specimen's amplicon with sequencing errors, extra flank on either side
                s = rnd(rng, rng.randint(0, 40)) + mutate(rng, amplicon, rng.choice([0.03, 0.07, 0.12, 0.16, 0.2])) + rnd(rng, rng.randint(0, 40))
            elif r < 0.6:      # truncated copy: the full read cannot be placed inside it within the threshold
                s = mutate(rng, amplicon[: int(amplicon_len * rng.uniform(0.5, 0.95))], 0.05)
            elif r < 0.8:      # another amplicon
                s = mutate(rng, other, 0.05)
            elif r < 0.9:      # chimera
                s = amplicon[: amplicon_len // 2] + other[amplicon_len // 2:]
            else:
                s = rnd(rng, rng.randint(30, amplicon_len))
            if i % 11 == 3:
                s = s[:50] + "N" + s[51:]
            recs.append(("p%s%02d\tdx:i:%d extra words" % (kind, i, i), s, qual(s)))
        partials[kind] = recs
    files = {"index.txt": "SampleID\tPrimerPool\tFwIndex\tFwPrimer\tRvIndex\tRvPrimer\nSP0\tITS\t%s\tITS1F\t%s\tITS4\nSP1\tITS\t%s\tITS1F\t%s\tITS4\n"
                          % (rnd(rng, 13), rnd(rng, 13), b1.lower(), b2)}
    if level == "pool":
        files["out/full/ITS/SP1.fastq"] = fastq(full)
        files["out/partial/ITS/ITS1F-ITS4/barcode_fwd_%s.fastq" % b1] = fastq(partials["fwd"][: n_partial // 2])
        files["out/partial/ITS/ITS1F-unknown/barcode_fwd_%s.fastq" % b1] = fastq(partials["fwd"][n_partial // 2:])
        files["out/partial/ITS/ITS1F-ITS4/barcode_rev_%s.fastq" % b2] = fastq(partials["rev"])
        target = "out/full/ITS/SP1.fastq"
    else:
        files["out/full/ITS/ITS1F-ITS4/SP1.fastq"] = fastq(full)
        files["out/partial/ITS/ITS1F-ITS4/barcode_fwd_%s.fastq" % b1] = fastq(partials["fwd"])
        files["out/partial/ITS/ITS1F-ITS4/sample_barcode_rev_%s.fastq" % b2] = fastq(partials["rev"])     # legacy name
        files["out/partial/ITS/ITS1F-unknown/barcode_rev_%s.fastq" % b2] = fastq(partials["rev"][:3])     # not looked at
        target = "out/full/ITS/ITS1F-ITS4/SP1.fastq"
    return files, target


def run_reference(files, target, flags):
    d = tempfile.mkdtemp(prefix="smx_mine_")
    try:
        for rel, text in files.items():
            p = os.path.join(d, rel)
            os.makedirs(os.path.dirname(p), exist_ok=True)
            with open(p, "w") as fh:
                fh.write(text)
        argv = ["specimine", "--index", os.path.join(d, "index.txt"), "--fastq", os.path.join(d, target)] + flags
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(HERE, "standins"), os.path.join(REF, "src"), ROOT]),
                   PYTHONHASHSEED="0")
        code = "import sys; sys.argv=%r; from specimux.cli import specimine_main; specimine_main()" % (argv,)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=d)
        mined = os.path.join(d, target + ".mined")
        return r.returncode, (open(mined).read() if os.path.exists(mined) else None), r.stderr
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    cases = []
    specs = [(1, 320, 6, 18, "pool", ["--partial-forward"]), (2, 180, 5, 14, "pair", []),
             (3, 700, 4, 12, "pool", ["--min-identity", "0.9", "--partial-forward", "--no-partial-reverse"]),
             (4, 1100, 3, 8, "pair", ["--partial-forward", "--min-identity", "0.8"]),
             (5, 90, 8, 20, "pool", ["--min-identity", "0.75"])]
    for seed, alen, nf, np_, level, flags in specs:
        files, target = make_case(seed, alen, nf, np_, level)
        rc, mined, err = run_reference(files, target, flags)
        assert rc == 0 and mined is not None, err[-2000:]
        n = mined.count("\n") // 4
        print("case seed=%d len=%d level=%s flags=%s: %d mined records" % (seed, alen, level, flags, n))
        assert n > 0
        cases.append({"files": files, "target": target, "flags": flags, "mined": mined})
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with gzip.open(OUT, "wt") as fh:
        json.dump(cases, fh)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
