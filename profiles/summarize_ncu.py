#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` output into a small markdown table (one column per launch).
usage: summarize_ncu.py raw.csv > summary.md"""
import csv
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    launches = [r for r in rows[hdr + 2:] if len(r) == len(names)]
    print("| metric | " + " | ".join("launch %d" % (i + 1) for i in range(len(launches))) + " | unit |")
    print("|---|" + "---|" * (len(launches) + 1))
    for k in KEYS:
        if k not in names:
            continue
        j = names.index(k)
        vals = [r[j][:48] for r in launches]
        print("| %s | %s | %s |" % (k, " | ".join(vals), units[j]))


if __name__ == "__main__":
    main(sys.argv[1])
