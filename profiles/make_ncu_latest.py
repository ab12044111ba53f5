#!/usr/bin/env python3
"""raw.csv (ncu -i X.ncu-rep --page raw --csv) -> profiles/ncu_latest.json: per stage kernel of ONE step, the
figures bench.py quotes beside its live timings (executed ALU-pipe utilisation, DRAM bytes per launch).
usage: make_ncu_latest.py raw.csv "capture description" > profiles/ncu_latest.json"""
import csv
import json
import sys

NAMES = [("k_stage_windows", "stage_windows"), ("k_primer_sliced", "primer_sliced"), ("k_primer_search", "primer_finish"),
         ("k_primer_start", "primer_start"), ("k_barcode_bitsliced", "barcode_bitsliced"), ("k_select_fast", "select_fast"),
         ("k_select<", "select_general"), ("k_scan_compact", "scan_compact"), ("k_rebase_offsets", "rebase_offsets"), ("k_scan", "scan"),
         ("k_compact_records", "compact_records")]


def main(path, capture):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    out = {}
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        kname = r[col["Kernel Name"]]
        key = next((k for pat, k in NAMES if pat in kname), None)
        if key is None:
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
    for d in out.values():
        d["alu_pipe_pct"] = sum(d.pop("_alu")) / d["launches"]
        d["issue_active_pct"] = sum(d.pop("_issue")) / d["launches"]
    json.dump({"capture": capture, "kernels": out}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
