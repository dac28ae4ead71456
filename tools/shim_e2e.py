#!/usr/bin/env python3
"""Wall-clock of the reference-facing C++ shims next to the reference's own code, through
oracle/_ref/shim_harness on 2,000 sprite frames: frc::collector vs frc_b200::collector (gpu_blit), fdf::filter vs
fdf_b200::filter (upload mode and resident mode).  The harness prints and checks identity itself."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from remap_b200 import synth  # noqa: E402


def main():
    seq = synth.scrolling_tilemap(2000, 320, 224, seed=3, sprites=8)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "f.bin")
        seq.frames.tofile(p)
        rc = 0
        for mode in ("1", "2"):  # 1: compatible collector (compressed copies kept); 2: lean (frames stay on the device)
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "shim_harness"), p, "320", "224", "2000", "512", "0", mode, "1", "0"],
                               capture_output=True, text=True)
            print(f"gpu_blit={mode}:", r.stdout[-800:], r.stderr[-300:])
            rc |= r.returncode
        return rc


if __name__ == "__main__":
    sys.exit(main())
