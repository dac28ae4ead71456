"""CPU, build container only: the C restatement against the REAL reference run live
(oracle/_ref/ref_harness) on larger seeded sequences than the committed fixtures.  Skipped where
oracle/_ref was never built (it needs /root/reference)."""
import numpy as np
import pytest

import parity
from oracle import oracle, refdump
from remap_b200 import synth

pytestmark = pytest.mark.skipif(not refdump.have_ref(), reason="oracle/_ref/ref_harness not built")

CASES = {
    "S": lambda: synth.scrolling_tilemap(24, 320, 224, seed=21).frames,
    "S_dense": lambda: synth.scrolling_tilemap(10, 320, 224, seed=22, speckle=0.10, detail=3).frames,
    "S_repeat": lambda: synth.scrolling_tilemap(10, 320, 224, seed=23, speckle=0.10, n_tiles=4).frames,
    "L": lambda: synth.scrolling_tilemap(5, 640, 480, seed=24, speckle=0.10, vmax=(48, 48)).frames,
    "cuts": lambda: synth.scrolling_tilemap(24, 320, 224, seed=25, cut_every=6, levels=3).frames,
    "sprites": lambda: synth.scrolling_tilemap(16, 320, 224, seed=26, sprites=12).frames,
    "odd": lambda: synth.scrolling_tilemap(12, 323, 227, seed=27).frames,
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_vs_live_reference(name):
    frames = CASES[name]()
    ref = refdump.ref_dump(frames)
    N, H, W = frames.shape
    cfg = oracle.config(W, H)
    prev = None
    for i in range(N):
        med, kps = oracle.extract(cfg, frames[i])
        parity.check_frame(ref["frames"][i], med, kps, f"{name} frame {i}")
        if i > 0:
            res, votes = oracle.match(cfg, prev, kps)
            bins = [oracle.region_bins(cfg, prev, kps, r) for r in range(8)]
            assert parity.check_pair(ref["pairs"][i - 1], res, votes, bins, f"{name} pair {i}") == "ok"
        prev = kps


def test_ground_truth_offsets():
    seq = synth.scrolling_tilemap(40, 320, 224, seed=28)
    out = oracle.register(oracle.config(320, 224), seq.frames)
    assert out["results"]["valid"].all()
    assert np.array_equal(np.stack([out["results"]["dx"], out["results"]["dy"]], 1), seq.true_offsets)
