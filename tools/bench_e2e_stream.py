#!/usr/bin/env python3
"""e2e experiments on rb_register_host_async: chunk size x lane x context options; prints every repetition.
usage: bench_e2e_stream.py <lane auto|packed|raw> [profile 0|1] [torch_stream 0|1] [chunks...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import remap_b200  # noqa: E402
from remap_b200 import synth  # noqa: E402

n, W, H = 20000, 320, 224
seq = synth.scrolling_tilemap(n, W, H, seed=1)
pinned = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
pinned.numpy()[...] = seq.frames
host = pinned.numpy()
out = np.zeros(n - 1, remap_b200.OFFSET_DTYPE)
lane = sys.argv[1]
profile = len(sys.argv) > 2 and sys.argv[2] == "1"
tstream = len(sys.argv) > 3 and sys.argv[3] == "1"
chunks = [int(x) for x in sys.argv[4:]] or [512, 1024]
if lane != "auto":
    os.environ["RB_HOST_LANE"] = lane
for chunk in chunks:
    st = torch.cuda.Stream() if tstream else None
    with remap_b200.Registrar(W, H, max_frames=n, upload_chunk=chunk, profile=profile,
                              stream=st.cuda_stream if st else None) as reg:
        ts = []
        for rep in range(6):
            t0 = time.perf_counter()
            reg.register_host_async(host)
            reg.fetch_offsets(n - 1, out=out)
            ts.append(round((time.perf_counter() - t0) * 1e3, 2))
        stt = reg.host_lane_stats
    print(json.dumps(dict(lane=lane, chunk=chunk, profile=profile, torch_stream=tstream, ms=ts, fps=round(n / (min(ts[2:]) * 1e-3)), **stt)), flush=True)
