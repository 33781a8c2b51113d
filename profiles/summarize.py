#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` dumps into one markdown table per capture.
usage: summarize.py label=raw.csv [label=raw.csv ...] > r1_summary.md"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("smsp__inst_executed.sum", "warp inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_sel")]

for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("### %s\n" % label)
    print("| kernel | grid×block | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue  # first launch of each kernel
        seen.add(name)
        cells = []
        for k, _ in KEYS:
            if k not in idx:
                cells.append("-")
                continue
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v)
                v = ("%.3g" % f) if abs(f) < 1e6 else ("%.4g" % f)
            except ValueError:
                pass
            cells.append((v + " " + u).strip())
        print("| %s | %s×%s | %s |" % (name, r[idx["Grid Size"]] if "Grid Size" in idx else r[idx["launch__grid_size"]], r[idx["Block Size"]] if "Block Size" in idx else r[idx["launch__block_size"]], " | ".join(cells)))
    print()
