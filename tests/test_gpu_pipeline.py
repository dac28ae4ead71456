"""GPU: the reference's WHOLE pipeline (mpb::builder::build: aws::scan -> frc -> fgs -> fdf -> arf, src/mpb.hpp:28-41)
once unchanged and once with the three substitutions of INTEGRATION.md (frc_b200::collector, fgs_b200::splice,
fdf_b200::filter), through oracle/_ref/pipeline_harness: the fragments handed over after every stage and the
final maps must be identical.  BASELINE configs[0] is this pipeline on a 320x224 window."""
import os
import subprocess

import numpy as np
import pytest

from remap_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PIPE = os.path.join(ROOT, "oracle", "_ref", "pipeline_harness")


def screen_sequence(n, seed, **kw):
    """A 388x312 screen (src/main.cpp:198) with a static border around a 326x230 changing area: aws::scan finds
    the contour and window_info shrinks it to the 323x227 action window (src/aws.hpp:74-83)."""
    seq = synth.scrolling_tilemap(n, 326, 230, seed=seed, world_w=1024, world_h=768, **kw)
    screen = np.full((n, 312, 388), 6, np.uint8)
    screen[:, 40:270, 30:356] = seq.frames
    return screen


def run(screen, tmp_path, mode="both"):
    if not os.path.exists(PIPE):
        pytest.fail("oracle/_ref/pipeline_harness missing: run `python oracle/build_ref.py` in the build container")
    n, H, W = screen.shape
    path = os.path.join(tmp_path, "screen.bin")
    np.ascontiguousarray(screen, np.uint8).tofile(path)
    r = subprocess.run([PIPE, path, str(W), str(H), str(n), mode], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, (r.stdout[-800:], r.stderr[-500:])
    assert ("PIPELINE IDENTICAL" if mode == "both" else "FAST PIPELINE IDENTICAL") in r.stdout, r.stdout
    return r.stdout


def test_whole_pipeline_one_fragment(tmp_path):
    out = run(screen_sequence(260, 81), str(tmp_path))
    assert "1 final map(s)" in out and "frc: 1 fragment(s)" in out, out


def test_whole_pipeline_sprites_and_cuts(tmp_path):
    """Scene cuts open several fragments (spliced back where they overlap), sprites exercise pass 2."""
    out = run(screen_sequence(400, 82, sprites=5, cut_every=90), str(tmp_path))
    print(out)
    assert "frc:" in out and "fdf:" in out, out


def test_fast_builder_equals_reference_pipeline(tmp_path):
    """include/mpb_b200.hpp (one resident collector, splice on its fragments, pass 2 in place) against
    mpb::builder::build: same fragments after every stage, same final maps."""
    out = run(screen_sequence(400, 83, sprites=4, cut_every=120), str(tmp_path), mode="fast")
    print(out)
