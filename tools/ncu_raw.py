#!/usr/bin/env python3
"""Pick the metrics that matter out of `ncu -i X.ncu-rep --page raw --csv`.
usage: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_raw.py [regex]"""
import csv
import re
import sys

pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else
                 r"gpu__time_duration.sum|sm__inst_executed.sum$|smsp__issue_active.avg.pct|pipe_alu.avg.pct_of_peak_sustained_active|"
                 r"pipe_lsu.avg.pct_of_peak_sustained_active|bank_conflicts_pipe_lsu_mem_shared.sum|wavefronts_mem_shared.sum$|"
                 r"thread_inst_executed_per_inst_executed.ratio|warps_active.avg.pct|dram__bytes_(read|write).sum$|registers_per_thread|"
                 r"dram__throughput.avg.pct|l1tex__data_pipe_lsu_wavefronts.avg.pct|smsp__warp_issue_stalled.*_per_warp_active.pct|"
                 r"smsp__average_warp.*_per_issue_active|lsu_mem_shared_op_(ld|st|atom).sum$|sm__throughput.avg.pct")
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f"  {h} = {r[i]} {units[i]}")
