#!/usr/bin/env python3
"""Static SASS opcode histogram of the stage kernels in the built objects (cuobjdump -sass), as a markdown table:
how much of each kernel's instruction stream is three-input logic (LOP3, the DP cells), moves on the FMA pipe
(IMAD.MOV), shared / global memory, and control.  usage: python tools/sass_hist.py [pattern ...] > profiles/rN_sass.md"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ["k_stage_windows", "k_primer_slicedILi22E", "k_primer_slicedILi20E", "k_primer_finishIjE",
           "k_barcode_taskILi3ELi1ELi13E", "k_barcode_taskILi3ELi3ELi13E", "k_select_fastILi8E", "k_selectILi2E",
           "k_scan_compact"]
GROUPS = [("LOP3", ("LOP3", "PLOP3", "ULOP3")), ("IMAD (FMA pipe; mostly moves)", ("IMAD", "UIMAD")),
          ("shift / funnel (SHF, LEA, PRMT, BREV, POPC, FLO)", ("SHF", "LEA", "PRMT", "BREV", "POPC", "FLO", "USHF", "ULEA")),
          ("integer add / compare / select", ("IADD3", "VIADD", "UIADD3", "ISETP", "UISETP", "SEL", "USEL", "VIMNMX", "VIADDMNMX",
                                             "VIMNMX3", "IABS", "LOP")),
          ("shared memory (LDS, STS)", ("LDS", "STS", "LDSM")), ("global / local memory", ("LDG", "STG", "LDL", "STL", "ATOM", "ATOMG", "RED", "ATOMS")),
          ("constant / uniform loads", ("LDC", "LDCU", "S2R", "S2UR", "CS2R", "UMOV", "MOV", "R2UR")),
          ("control (BRA, BSSY, BSYNC, BAR, EXIT, ...)", ("BRA", "BSSY", "BSYNC", "BAR", "EXIT", "BREAK", "WARPSYNC", "NOP", "CALL", "RET")),
          ("warp (SHFL, VOTE, REDUX, MATCH)", ("SHFL", "VOTE", "REDUX", "MATCH"))]


def main(patterns):
    rows = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "specimux_b200", "csrc", "build", "*.o"))):
        if re.search(r"stage2_k(?!3\.o)", obj):
            continue                                  # one threshold's stage-2 object is enough (K = 3: config 2)
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        fn, ops = None, None
        for line in out.splitlines():
            m = re.match(r"\s+Function : (\S+)", line)
            if m:
                if fn and ops:
                    rows.append((fn, ops))
                fn = m.group(1) if any(p in m.group(1) for p in patterns) else None
                ops = collections.Counter()
                continue
            if fn:
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
                if m:
                    ops[m.group(1)] += 1
        if fn and ops:
            rows.append((fn, ops))
    print("| kernel (mangled) | instructions | " + " | ".join(g for g, _ in GROUPS) + " | other |")
    print("|---|---|" + "---|" * (len(GROUPS) + 1))
    for fn, ops in rows:
        tot = sum(ops.values())
        cells, seen = [], 0
        for _g, names in GROUPS:
            n = sum(ops[x] for x in names)
            seen += n
            cells.append("%d (%.0f %%)" % (n, 100.0 * n / tot))
        print("| `%s` | %d | %s | %d |" % (fn[:60], tot, " | ".join(cells), tot - seen))


if __name__ == "__main__":
    main(sys.argv[1:] or DEFAULT)
