#!/usr/bin/env python3
"""raw.csv (ncu -i X.ncu-rep --page raw --csv) -> profiles/ncu_latest.json: per stage kernel of ONE step, the
figures bench.py quotes beside its live timings (executed ALU-pipe utilisation, DRAM bytes per launch).
The json carries the sha1 prefix of the library SOURCES the capture was taken from (`build_id`, _lib.source_id()); bench.py quotes these figures
only when it is running a library built from those very sources.
usage: make_ncu_latest.py raw.csv "capture description" [path/to/libspecimux_b200.so] > profiles/ncu_latest.json"""
import csv
import hashlib
import json
import os
import sys

NAMES = [("k_stage_windows", "stage_windows"), ("k_primer_sliced", "primer_sliced"), ("k_primer_finish", "primer_finish_start"),
         ("k_primer_long", "primer_long"), ("k_barcode_task", "barcode_tasks"), ("k_select_fast", "select_fast"),
         ("k_select<", "select_general"), ("k_scan_compact", "scan_compact"), ("k_rebase_offsets", "rebase_offsets")]


def main(path, capture, lib_path=None):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    out = {}
    seen_stage = False
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        kname = r[col["Kernel Name"]]
        key = next((k for pat, k in NAMES if pat in kname), None)
        if key is None:
            continue
        if key == "stage_windows":          # ONE step: from the first window staging to just before the next one
            if seen_stage:
                break
            seen_stage = True
        if not seen_stage:
            continue

        def val(metric, scale_unit=True):
            v = float(r[col[metric]].replace(",", ""))
            u = units[col[metric]]
            if scale_unit:
                v *= {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
            return v
        d = out.setdefault(key, {"launches": 0, "duration_us": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0, "_alu": [], "_issue": []})
        d["launches"] += 1
        d["duration_us"] += val("gpu__time_duration.sum")
        d["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        d["warp_inst"] += val("smsp__inst_executed.sum")
        d["_alu"].append(val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"))
        d["_issue"].append(val("smsp__issue_active.avg.pct_of_peak_sustained_active"))
    t_all = a_all = 0.0
    for d in out.values():
        d["alu_pipe_pct"] = sum(d.pop("_alu")) / d["launches"]
        d["issue_active_pct"] = sum(d.pop("_issue")) / d["launches"]
        t_all += d["duration_us"]
        a_all += d["duration_us"] * d["alu_pipe_pct"]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from specimux_b200 import _lib
    build_id = _lib.source_id()             # sources of the library the capture ran (bench.py compares the same id)
    json.dump({"capture": capture, "build_id": build_id, "step_us": t_all,
               "whole_step_alu_pipe_pct": a_all / t_all if t_all else None,
               "whole_step_dram_bytes": sum(d["dram_bytes"] for d in out.values()), "kernels": out}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", sys.argv[3] if len(sys.argv) > 3 else None)
