#!/usr/bin/env python
"""Summarise an ncu report (read on the CPU box): python profiles/ncu_summary.py file.ncu-rep [regex]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|launch__(registers_per_thread|grid_size|block_size|occupancy_limit)|"
                 r"sm__warps_active.avg.pct|smsp__issue_active.avg.pct|sm__inst_executed_pipe_(fma|alu|xu|lsu|fp64)\.avg.pct|sm__pipe_fp64_cycles_active.avg.pct|"
                 r"smsp__inst_executed.sum$|smsp__warp_issue_stalled_.*_per_warp_active.pct|sm__throughput.avg.pct|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__thread_inst_executed_per_inst_executed.ratio|gpu__dram_throughput.avg.pct")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if pat.search(h):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "stalled" in h and v < 3.0:
                continue
            print(f"  {h:90s} {units[i]:12s} {r[i]}")
