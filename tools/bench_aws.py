#!/usr/bin/env python3
"""rb_aws_compare (aws::details::compare over a run of resident frames) on 2,000 screen-sized frames: wall clock
of the call (with its small host copies); the kernel time comes from ncu (profiles/README.md)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import remap_b200  # noqa: E402
from remap_b200 import synth  # noqa: E402


def main():
    n = 2000
    seq = synth.scrolling_tilemap(n, 320, 224, seed=8)
    screen = np.full((n, 312, 388), 14, np.uint8)  # the reference's screen size (src/main.cpp:194-244)
    screen[:, 40:264, 32:352] = seq.frames
    with remap_b200.Registrar(388, 312, max_frames=n) as reg:
        reg.upload(screen)
        ts = []
        for _ in range(6):
            t = time.perf_counter()
            reg.aws_compare(n)
            ts.append(time.perf_counter() - t)
    print(json.dumps(dict(frames=n, wall_ms=min(ts) * 1e3, bytes=int(n * 388 * 312), GBps_wall=n * 388 * 312 / min(ts) / 1e9)))


if __name__ == "__main__":
    main()
