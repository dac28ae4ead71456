import sys, time, json, numpy as np
sys.path.insert(0,'/root/repo')
import remap_b200, ctypes as C
from remap_b200 import synth
n=2000
seq = synth.scrolling_tilemap(n, 320, 224, seed=8)
screen = np.full((n, 312, 388), 14, np.uint8); screen[:, 40:264, 32:352] = seq.frames
with remap_b200.Registrar(388, 312, max_frames=n) as reg:
    reg.upload(screen)
    ts=[]
    import torch
    for r in range(6):
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        t=time.perf_counter(); heat, fc = reg.aws_compare(n); ts.append(time.perf_counter()-t)
    print(json.dumps(dict(frames=n, wall_ms=min(ts)*1e3, bytes=int(n*388*312), GBps_wall=n*400*312/min(ts)/1e9)))
