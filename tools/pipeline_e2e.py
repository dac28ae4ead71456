#!/usr/bin/env python3
"""BASELINE configs[0]: the reference's whole CPU pipeline (mpb::builder::build) on a synthetic screen sequence,
next to the same pipeline with the three B200 shims, through oracle/_ref/pipeline_harness (identity checked there)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remap_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    seq = synth.scrolling_tilemap(n, 326, 230, seed=1, world_w=4096, world_h=2048)
    screen = np.full((n, 312, 388), 6, np.uint8)
    screen[:, 40:270, 30:356] = seq.frames
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "s.bin")
        screen.tofile(p)
        rc = 0
        for mode in ("both", "fast", "ref", "b200", "fast-time"):  # the last three: one pipeline per process, for timing
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "pipeline_harness"), p, "388", "312", str(n), mode],
                               capture_output=True, text=True)
            print(r.stdout[-600:], r.stderr[-300:])
            rc |= r.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main())
