import sys, os, subprocess, tempfile, numpy as np
sys.path.insert(0, '/root/repo')
from remap_b200 import synth
seq = synth.scrolling_tilemap(2000, 320, 224, seed=3, sprites=8)
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, 'f.bin'); seq.frames.tofile(p)
    r = subprocess.run(['/root/repo/oracle/_ref/shim_harness', p, '320', '224', '2000', '512', '0', '1', '1', '0'], capture_output=True, text=True)
    print(r.stdout[-600:], r.stderr[-300:])
