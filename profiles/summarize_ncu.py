#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/.
usage: python profiles/summarize_ncu.py raw.csv out.csv"""
import csv
import re
import sys

PATTERNS = [
    r"^ID$", r"^Kernel Name$", r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__throughput\.avg\.pct", r"^lts__throughput\.avg\.pct",
    r"^smsp__issue_active\.avg\.pct", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^launch__registers_per_thread$", r"^launch__occupancy_limit_(registers|shared_mem|warps)$", r"^launch__shared_mem_per_block",
    r"^launch__grid_size$", r"^launch__block_size$",
    r"^l1tex__t_(sectors|requests)_pipe_lsu_mem_global_op_(ld|st)\.sum$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
    r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$", r"^lts__t_sector_hit_rate\.pct$",
    r"^sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|lsu|xu|uniform)\.sum$", r"^sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|lsu|xu)\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio$",
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    h = rows[0]
    idx = [i for i, name in enumerate(h) if any(re.search(p, name) for p in PATTERNS)]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    # human-readable transposed view
    for r in rows[2:]:
        print("=" * 100)
        for i in idx:
            v = r[i]
            if h[i].startswith("smsp__pcsamp") and v in ("0", "0.000000"):
                continue
            print(f"{h[i][:78]:80s}{v[:60]} {rows[1][i]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
