#!/usr/bin/env python3
"""BASELINE configs[3] in small: 640x480, large scroll offsets, dense keypoints (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import remap_b200  # noqa: E402
from remap_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seq = synth.scrolling_tilemap(n, 640, 480, seed=4, speckle=0.10, vmax=(48, 36), world_w=8192, world_h=4096)
with remap_b200.Registrar(640, 480, max_frames=n, profile=True) as reg:
    reg.upload(seq.frames)
    for _ in range(2):
        off, _ = reg.register(n)
    print(reg.kernel_times(), "deferred", reg.deferred_count)
