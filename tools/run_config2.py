#!/usr/bin/env python3
"""BASELINE configs[1] in small (for ncu captures): 320x224, n frames, two registrations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import remap_b200  # noqa: E402
from remap_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
seq = synth.scrolling_tilemap(n, 320, 224, seed=1)
with remap_b200.Registrar(320, 224, max_frames=n, profile=True) as reg:
    reg.upload(seq.frames)
    for _ in range(2):
        off, _ = reg.register(n)
    print(reg.kernel_times(), "deferred", reg.deferred_count)
