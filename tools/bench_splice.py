#!/usr/bin/env python3
"""fgs::splice (reference, one host thread) next to fgs_b200::splice (B200) on the reference collector's own
fragments, through oracle/_ref/shim_harness; plus rb_snippet_create / rb_snippet_match timings on large maps.
Prints one JSON object; not part of bench.py's contract line."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import remap_b200  # noqa: E402
from remap_b200 import synth  # noqa: E402

SHIM = os.path.join(ROOT, "oracle", "_ref", "shim_harness")


def main():
    res = {}
    seq = synth.scrolling_tilemap(n=600, w=320, h=224, seed=43, world_w=1280, world_h=896, cut_every=60)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "f.bin")
        seq.frames.tofile(path)
        r = subprocess.run([SHIM, path, "320", "224", "600", "64", "0", "0", "0", "1"], capture_output=True, text=True, timeout=1800)
    line = [l for l in r.stdout.splitlines() if l.startswith("SPLICE")]
    res["shim_harness"] = line[0] if line else r.stdout[-300:]
    m = re.search(r"fgs::splice ([0-9.]+) ms, fgs_b200::splice ([0-9.]+) ms", res["shim_harness"])
    if m:
        res["reference_ms"], res["b200_ms"] = float(m.group(1)), float(m.group(2))
    tr = [l for l in r.stderr.splitlines() if l.startswith("fgs_b200::splice:")]
    if tr:
        res["phases"] = tr[-1]  # RB_SPLICE_TRACE=1
    # large maps: two overlapping 2400x1600 crops of one world
    rng = np.random.default_rng(7)
    world = synth.make_world(rng, 4096, 2048, n_tiles=64, speckle=0.05)

    def dots_of(img):
        d = np.zeros(img.shape + (16,), np.uint16)
        np.put_along_axis(d, img[:, :, None].astype(np.int64), 2, axis=2)
        return d

    a, b = dots_of(world[100:1700, 200:2600]), dots_of(world[300:1900, 900:3300])
    t0 = time.perf_counter()
    sa = remap_b200.Snippet(a)
    t1 = time.perf_counter()
    sb = remap_b200.Snippet(b)
    ka = sa.fetch()["kps"]
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        mm = sa.match(sb)
        ts.append(time.perf_counter() - t)
    res["large"] = dict(map=[2400, 1600], keypoints=int(len(ka)), create_ms=round((t1 - t0) * 1e3, 1),
                        match_ms=round(min(ts) * 1e3, 2), match={k: int(mm[k]) for k in mm.dtype.names})
    sa.close(); sb.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
