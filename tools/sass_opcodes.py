#!/usr/bin/env python3
"""Opcode counts per kernel from `cuobjdump -sass libremap_b200.so` (evidence of what the kernels are made of:
UTMALDG / UBLKCP = TMA tensor / bulk copies, SYNCS = mbarrier, REDUX / MATCH = warp reductions, LOP3 = the bit-sliced logic).
usage: python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "remap_b200", "libremap_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        counts[cur][m.group(1)] += 1
keys = ["UTMALDG", "UBLKCP", "SYNCS", "REDUX", "MATCH", "VOTE", "SHFL", "LOP3", "PRMT", "SHF", "IMAD", "ATOMS", "LDS", "STS", "LDG", "STG", "BAR"]
print("kernel".ljust(34), "total", " ".join(k.rjust(7) for k in keys))
for name, c in counts.items():
    print(name[:34].ljust(34), str(sum(c.values())).rjust(5), " ".join(str(c.get(k, 0)).rjust(7) for k in keys))
