import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import remap_b200
from remap_b200 import synth
n=20000; W,H=320,224
seq = synth.scrolling_tilemap(n, W, H, seed=1)
pinned = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True); pinned.numpy()[...] = seq.frames
host = pinned.numpy(); out = np.zeros(n - 1, remap_b200.OFFSET_DTYPE)
lane = sys.argv[1]
if lane != "auto": os.environ["RB_HOST_LANE"] = lane
for chunk in (128, 256, 512, 1024):
    with remap_b200.Registrar(W, H, max_frames=n, upload_chunk=chunk) as reg:
        best = None
        for rep in range(5):
            t0 = time.perf_counter(); reg.register_host_async(host); reg.fetch_offsets(n - 1, out=out); dt = time.perf_counter() - t0
            if rep >= 2 and (best is None or dt < best): best = dt
        st = reg.host_lane_stats
    print(json.dumps(dict(stream=os.environ.get("RB_PACK_STREAM"), lane=lane, chunk=chunk, fps=n / best, **st)), flush=True)
