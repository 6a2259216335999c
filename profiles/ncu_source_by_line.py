#!/usr/bin/env python
"""Instructions executed and stall samples per SOURCE LINE of a kernel in an ncu report captured with
--import-source on (kernel compiled with -lineinfo): python profiles/ncu_source_by_line.py file.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0, ""])
tot_i = tot_s = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] in ("Function Name",) or not r[0].strip().isdigit():
        continue
    try:
        n, s = int(r[ii]), int(r[si])
    except (ValueError, IndexError):
        continue
    key = (cur_file, int(r[0]))
    agg[key][0] += n
    agg[key][1] += s
    agg[key][2] = r[1].strip()[:90]
    tot_i += n
    tot_s += s
print(f"total instructions {tot_i}, samples {tot_s}")
by_file = defaultdict(lambda: [0, 0])
for (f, _), v in agg.items():
    by_file[f][0] += v[0]
    by_file[f][1] += v[1]
for f, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:22s} instr {v[0]:>10d} ({100.0 * v[0] / tot_i:5.1f} %)  samples {v[1]:>6d} ({100.0 * v[1] / max(1, tot_s):5.1f} %)")
print("top lines by instructions:")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {f}:{ln:<5d} {v[0]:>9d} ({100.0 * v[0] / tot_i:4.1f} %) samples {v[1]:>5d} ({100.0 * v[1] / max(1, tot_s):4.1f} %)  {v[2]}")
