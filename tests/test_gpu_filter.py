"""GPU: pass-2 foreground filtering (rb_filter_fragment = fdf::filter, src/fdf.hpp:40-75) through the C ABI
against (a) committed dumps of the REAL reference's frc::collector + fdf::filter and (b) the C/numpy
restatement on larger seeded sequences; shared-memory variant, forced deferrals and the general variant."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

import remap_b200
from oracle import oracle, refdump
from remap_b200 import synth
from remap_b200.api import PLACEMENT_DTYPE

pytestmark = pytest.mark.gpu

FILTER_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "filter_*.npz")))


def places_of(idx, pos):
    pl = np.zeros(len(idx), PLACEMENT_DTYPE)
    pl["frame"], pl["x"], pl["y"] = idx, pos[:, 0], pos[:, 1]
    return pl


@pytest.mark.parametrize("name", FILTER_CASES)
def test_filter_matches_reference_dump(name, golden_dir):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    frames, ref = z["frames"], refdump.parse_filter_dump(z["dump"].tobytes())
    N, H, W = frames.shape
    with remap_b200.Registrar(W, H, max_frames=N) as reg:
        reg.upload(frames)
        reg.register(N)
        for fi, frag in enumerate(ref["fragments"]):
            recs = [r for r in ref["frames"] if r["fragment"] == fi]
            pl = places_of([r["number"] for r in recs], np.array([[r["x"], r["y"]] for r in recs]))
            mh, mw = ref["backgrounds"][fi]["image"].shape
            out = reg.filter_fragment(pl, mw, mh, want_fgmasks=True)   # background computed on the device
            for k, r in enumerate(recs):
                assert np.array_equal(out["fgmasks"][k], r["mask"]), f"{name} frame {r['number']}: fde::mask"
                assert out["ncontours"][k] == len(r["contours"]), f"{name} frame {r['number']}: contours"
            assert np.array_equal(out["dots"], frag["dots"]), f"{name} fragment {fi}: filtered dots"
            # and with the background handed in, as the 5-argument fdf::filter takes it
            out2 = reg.filter_fragment(pl, mw, mh, background=ref["backgrounds"][fi]["image"], want_fgmasks=True)
            assert np.array_equal(out2["dots"], frag["dots"]) and np.array_equal(out2["fgmasks"], out["fgmasks"])


def run_case(frames, pos, reg=None):
    """-> (device result, oracle result) for one fragment holding all frames at positions pos."""
    N, H, W = frames.shape
    mw, mh = int(pos[:, 0].max()) + W, int(pos[:, 1].max()) + H
    own = reg is None
    if own:
        reg = remap_b200.Registrar(W, H, max_frames=N)
    try:
        reg.upload(frames)
        _, med = reg.register(N, want_medians=True)
        out = reg.filter_fragment(places_of(np.arange(N), pos), mw, mh, want_fgmasks=True)
    finally:
        if own:
            reg.close()
    want = oracle.filter_fragment(frames, med, pos, mw, mh)
    return out, want


def check(out, want, name):
    n = len(want["masks"])
    for i in range(n):
        assert np.array_equal(out["fgmasks"][i], want["masks"][i]), f"{name} frame {i}: mask ({int((out['fgmasks'][i] != want['masks'][i]).sum())} px)"
    assert np.array_equal(out["ncontours"], want["ncontours"]), name
    assert np.array_equal(out["dots"], want["dots"]), f"{name}: dots"
    best = want["dots"].max(axis=2)
    assert np.array_equal(out["image"], np.where(best != 0, want["dots"].argmax(axis=2), 0))
    assert np.array_equal(out["mask"], (best != 0).astype(np.uint8))


def test_filter_sprites_320x224():
    seq = synth.scrolling_tilemap(40, 320, 224, seed=31, sprites=10, world_w=640, world_h=448)
    out, want = run_case(seq.frames, seq.path - seq.path.min(axis=0))
    check(out, want, "sprites")
    assert out["frames_deferred"] == 0


def test_filter_odd_size_and_640x480():
    seq = synth.scrolling_tilemap(6, 131, 99, seed=32, sprites=3, world_w=320, world_h=256)
    out, want = run_case(seq.frames, seq.path - seq.path.min(axis=0))
    check(out, want, "odd")
    seq = synth.scrolling_tilemap(4, 640, 480, seed=33, sprites=8, world_w=800, world_h=600)
    out, want = run_case(seq.frames, seq.path - seq.path.min(axis=0))
    check(out, want, "640x480")


def test_filter_random_frames_every_pixel_a_seed():
    frames = synth.random_frames(6, 96, 64, seed=34)
    pos = np.array([[0, 0], [3, 1], [5, 5], [1, 7], [9, 2], [4, 4]])
    out, want = run_case(frames, pos)
    check(out, want, "random")


def test_filter_flat_and_single_frame():
    frames = np.full((2, 64, 96), 3, np.uint8)
    frames[1, 10:20, 10:30] = 5
    out, want = run_case(frames, np.zeros((2, 2), np.int64))
    check(out, want, "flat")
    seq = synth.scrolling_tilemap(2, 160, 112, seed=35, world_w=320, world_h=256)
    out, want = run_case(seq.frames[:1].repeat(2, axis=0), np.zeros((2, 2), np.int64))
    check(out, want, "still")
    assert out["fgmasks"].sum() == 0   # the frame equals its own background: no seeds


CHILD = r"""
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import test_gpu_filter as T
from remap_b200 import synth
seq = synth.scrolling_tilemap(12, 320, 224, seed=36, sprites=8, world_w=640, world_h=448)
out, want = T.run_case(seq.frames, seq.path - seq.path.min(axis=0))
T.check(out, want, "child")
print("DEFERRED", out["frames_deferred"])
"""


@pytest.mark.parametrize("env,expect", [({"RB_FG_RCAP": "4096"}, 12), ({"RB_FG_SCAP": "16"}, 12),
                                        ({"RB_FG_GENERAL_ONLY": "1"}, 0), ({"RB_FG_CTAS": "1"}, 0)])
def test_filter_deferral_paths(env, expect):
    """Frames that do not fit the shared-memory tables take the general variant: same results."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", CHILD.format(root=root, tests=os.path.join(root, "tests"))], env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"DEFERRED {expect}" in r.stdout, r.stdout


def test_filter_needs_registered_frames():
    frames = synth.random_frames(3, 96, 64, seed=37)
    with remap_b200.Registrar(96, 64, max_frames=3) as reg:
        reg.upload(frames)
        with pytest.raises(remap_b200.RemapError):
            reg.filter_fragment(places_of([0], np.array([[0, 0]])), 96, 64)
        reg.register(3)
        with pytest.raises(remap_b200.RemapError):
            reg.filter_fragment(places_of([0], np.array([[1, 0]])), 96, 64)   # outside the map


def test_filter_edge_cases():
    """No placements (an empty fragment): all-zero dots, nothing masked.  One frame whose whole interior differs from
    the background: every contour seeded; the big ones dropped by the area rule."""
    frames = synth.random_frames(2, 96, 64, seed=38, palette=3)
    with remap_b200.Registrar(96, 64, max_frames=2) as reg:
        reg.upload(frames)
        _, med = reg.register(2, want_medians=True)
        out = reg.filter_fragment(np.zeros(0, PLACEMENT_DTYPE), 120, 80, want_fgmasks=True)
        assert out["dots"].sum() == 0 and out["image"].sum() == 0 and out["mask"].sum() == 0 and len(out["ncontours"]) == 0
        bg = np.full((80, 120), 15, np.uint8)       # a colour the 3-colour frames never use: every pixel a seed
        pl = places_of([1], np.array([[7, 9]]))
        out = reg.filter_fragment(pl, 120, 80, background=bg, want_fgmasks=True)
        want = oracle.filter_fragment(frames[1:2], med[1:2], np.array([[7, 9]]), 120, 80, background=bg)
        assert np.array_equal(out["fgmasks"], want["masks"]) and np.array_equal(out["dots"], want["dots"])
        assert np.array_equal(out["ncontours"], want["ncontours"])
