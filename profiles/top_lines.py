#!/usr/bin/env python3
"""Top CUDA source lines of one kernel by executed warp instructions and stall samples.
input: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:NAME > f.csv
usage: top_lines.py f.csv [N]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur, hdr, out = None, None, []
    for r in rows:
        if r and r[0] == "File Path":
            cur, hdr = r[1].split("/")[-1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if cur and hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":      # a CUDA line (SASS rows carry an address)
            ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
            try:
                out.append((int(r[ie]), int(r[ns] or 0), cur, r[0], r[1].strip()[:100]))
            except ValueError:
                pass
    tot, stot = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print("total warp instructions %d, stall samples %d" % (tot, stot))
    print("inst%  samp%  file:line  source")
    for n, s, f, l, src in sorted(out, reverse=True)[:top]:
        print("%5.1f  %5.1f  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / stot, f, l, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
