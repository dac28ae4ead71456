#!/usr/bin/env python3
"""Aggregate `ncu -i rep --page source --print-source cuda,sass --csv` by CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_by_line.py [top]"""
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or r[0] in ("Function Name",) or r[0] == "":
        continue
    try:
        out.append((cur, int(r[0]), int(r[ci]), int(r[si]), int(r[ti]), r[1]))
    except (ValueError, IndexError):
        pass
tot = sum(o[2] for o in out) or 1
ts = sum(o[3] for o in out) or 1
print(f"total warp-instructions {tot}, samples {ts}")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0]}:{o[1]:<4} samp {o[3] * 100 / ts:5.1f}%  inst {o[2] * 100 / tot:5.1f}%  lanes {o[4] / max(o[2], 1):5.1f}  {o[5].strip()[:100]}")
