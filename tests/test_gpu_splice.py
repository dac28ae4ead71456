"""GPU: fragment splicing through the C ABI: rb_snippet_create (fgs::details::extract_single)
and rb_snippet_match (the cellular kpm::match) against committed dumps of the REAL reference and the C
restatement; include/fgs_b200.hpp against the reference's fgs::splice through oracle/_ref/shim_harness."""
import glob
import os
import subprocess

import numpy as np
import pytest

import remap_b200
from oracle import oracle, refdump
from remap_b200 import synth
from test_oracle_golden import check_cell_match

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "shim_harness")
SPLICE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "splice_*.npz")))


@pytest.mark.parametrize("name", SPLICE_CASES)
def test_snippets_and_cell_matches_match_reference_dump(name, golden_dir):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    ref = refdump.parse_splice_dump(z["dump"].tobytes())
    snips = [remap_b200.Snippet(f["dots"]) for f in ref["fragments"]]
    try:
        for s, r, f in zip(snips, ref["snippets"], ref["fragments"]):
            got = s.fetch()
            img, msk = oracle.blend(f["dots"])
            assert np.array_equal(got["image"], img) and np.array_equal(got["mask"], r["mask"])
            k = got["kps"]
            order = np.lexsort((k["x"], k["y"]))
            for fld in ("x", "y", "code"):
                assert np.array_equal(k[fld][order], r["kps"][fld]), f"{name}: snippet keypoint field {fld}"
        for m in ref["matches"]:
            check_cell_match(snips[m["prev"]].match(snips[m["curr"]]), m, f"{name} {m['prev']}-{m['curr']}")
    finally:
        for s in snips:
            s.close()


def test_cell_match_large_maps_against_oracle():
    """Two overlapping 800x600 crops of one world (thousands of keypoints, repeated tiles): every field of the
    match record against the C restatement; plus a pair with nothing in common and an empty snippet."""
    rng = np.random.default_rng(61)
    world = synth.make_world(rng, 1280, 960, n_tiles=24, speckle=0.03)
    other = synth.make_world(rng, 800, 600, n_tiles=24, speckle=0.03)

    def dots_of(img, visits):
        d = np.zeros(img.shape + (16,), np.uint16)
        np.put_along_axis(d, img[:, :, None].astype(np.int64), visits, axis=2)
        return d

    a = dots_of(world[100:700, 200:1000], 3)
    b = dots_of(world[260:860, 410:1210], 2)
    b[:50] = 0                                           # part of the map never visited: mask 0, image 0
    c = dots_of(other, 1)
    e = np.zeros((96, 128, 16), np.uint16)
    dev = [remap_b200.Snippet(x) for x in (a, b, c, e)]
    ora = [oracle.snippet(x) for x in (a, b, c, e)]
    try:
        for i, j in ((0, 1), (1, 0), (0, 2), (2, 1), (0, 3), (3, 0)):
            got, want = dev[i].match(dev[j]), oracle.cell_match(ora[i], ora[j])
            for fld in ("offsets", "pairs", "ties", "matched_keypoints"):
                assert got[fld] == want[fld], (i, j, fld, got, want)
            if want["ties"] == 1:
                for fld in got.dtype.names:
                    assert got[fld] == want[fld], (i, j, fld, got, want)
        m = dev[0].match(dev[1])
        assert m["valid"] and (m["dx"], m["dy"]) == (210, 160)
    finally:
        for s in dev:
            s.close()


def _shim(frames, tmp_path):
    if not os.path.exists(SHIM):
        pytest.fail("oracle/_ref/shim_harness missing: run `python oracle/build_ref.py` in the build container")
    n, H, W = frames.shape
    path = os.path.join(tmp_path, "frames.bin")
    np.ascontiguousarray(frames, np.uint8).tofile(path)
    r = subprocess.run([SHIM, path, str(W), str(H), str(n), "32", "0", "0", "0", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-800:], r.stderr[-500:])
    return r.stdout


@pytest.mark.parametrize("kw,expect", [
    (dict(n=120, w=160, h=112, seed=41, world_w=400, world_h=304, cut_every=30), "3 fragments -> 1"),
    (dict(n=150, w=160, h=112, seed=42, world_w=360, world_h=264, cut_every=25, vmax=(3, 2)), "5 fragments -> 1"),
    (dict(n=60, w=128, h=96, seed=52, world_w=320, world_h=240, cut_every=15, levels=2), "5 fragments -> 3"),
    (dict(n=200, w=320, h=224, seed=43, world_w=800, world_h=608, cut_every=40), None),
])
def test_splice_shim_matches_reference_fgs_splice(kw, expect, tmp_path):
    """fgs_b200::splice against fgs::splice on the reference collector's own fragments: the same fragments come
    out, in the same order, dot for dot."""
    out = _shim(synth.scrolling_tilemap(**kw).frames, str(tmp_path))
    assert "SPLICE IDENTICAL" in out, out
    if expect:
        assert expect in out, out
