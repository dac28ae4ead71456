#!/usr/bin/env python3
"""e2e sweep of rb_register_host_async: lanes (auto / raw / packed), chunk size, packer threads.
usage: bench_e2e.py [frames] ; prints one JSON line per setting"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import remap_b200  # noqa: E402
from remap_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
W, H = 320, 224
seq = synth.scrolling_tilemap(n, W, H, seed=1)
pinned = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
pinned.numpy()[...] = seq.frames
host = pinned.numpy()
out = np.zeros(n - 1, remap_b200.OFFSET_DTYPE)
settings = []
for lane in ("packed", "raw", "auto"):
    for chunk in (256, 512, 1024, 2048):
        settings.append((lane, chunk, 0))
for th in (2, 4, 8):
    settings.append(("auto", 512, th))
    settings.append(("packed", 512, th))
for lane, chunk, th in settings:
    if lane == "auto":
        os.environ.pop("RB_HOST_LANE", None)
    else:
        os.environ["RB_HOST_LANE"] = lane
    with remap_b200.Registrar(W, H, max_frames=n, upload_chunk=chunk, host_threads=th) as reg:
        best = None
        for rep in range(5):
            t0 = time.perf_counter()
            reg.register_host_async(host)
            reg.fetch_offsets(n - 1, out=out)
            dt = time.perf_counter() - t0
            if rep >= 2 and (best is None or dt < best):
                best = dt
        st = reg.host_lane_stats
    ok = bool(np.array_equal(np.stack([out["dx"], out["dy"]], 1), seq.true_offsets))
    print(json.dumps(dict(lane=lane, chunk=chunk, cfg_threads=th, fps=n / best, ok=ok, **st)), flush=True)
