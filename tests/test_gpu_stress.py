"""GPU: seeded random stress of the rows next to the hot path against the C / numpy restatement -- frame sizes,
colour counts (few colours: large snaking components; many: thousands of tiny ones), seed densities and map
geometries that the structured workloads do not reach."""
import numpy as np
import pytest

import remap_b200
from oracle import oracle
from remap_b200.api import PLACEMENT_DTYPE

pytestmark = pytest.mark.gpu


def blobby(rng, n, H, W, colours, smooth):
    """Images with `colours` colours in patches of roughly `smooth` pixels (smooth = 1: pure noise)."""
    out = np.empty((n, H, W), np.uint8)
    for f in range(n):
        small = rng.integers(0, colours, size=((H + smooth - 1) // smooth + 1, (W + smooth - 1) // smooth + 1), dtype=np.uint8)
        big = np.kron(small, np.ones((smooth, smooth), np.uint8))
        oy, ox = rng.integers(0, smooth, size=2) if smooth > 1 else (0, 0)
        out[f] = big[oy:oy + H, ox:ox + W]
    return out


@pytest.mark.parametrize("case", range(10))
def test_filter_random_geometry(case):
    rng = np.random.default_rng(1000 + case)
    W = int(rng.integers(40, 400))
    H = int(rng.integers(40, 300))
    n = int(rng.integers(2, 7))
    colours = int(rng.choice([2, 3, 5, 16]))
    smooth = int(rng.choice([1, 2, 5, 13]))
    frames = blobby(rng, n, H, W, colours, smooth)
    change = rng.random(frames.shape) < rng.choice([0.002, 0.03, 0.4])     # what differs from the background: the seeds
    frames = np.where(change, rng.integers(0, 16, size=frames.shape, dtype=np.uint8), frames)
    mw, mh = W + int(rng.integers(0, 40)), H + int(rng.integers(0, 30))
    pos = np.stack([rng.integers(0, mw - W + 1, size=n), rng.integers(0, mh - H + 1, size=n)], 1)
    try:
        reg = remap_b200.Registrar(W, H, max_frames=n)
    except remap_b200.RemapError:
        pytest.skip(f"{W}x{H}: geometry outside the registration grid's limits")
    with reg:
        reg.upload(frames)
        _, med = reg.register(n, want_medians=True)
        pl = np.zeros(n, PLACEMENT_DTYPE)
        pl["frame"], pl["x"], pl["y"] = np.arange(n), pos[:, 0], pos[:, 1]
        # medians that are NOT kpe's: arbitrary images exercise run structures the rank filter never produces
        fake = blobby(rng, n, H, W, colours, smooth)
        for which, m in (("kpe", med), ("arbitrary", fake)):
            if which == "arbitrary":
                reg._check(reg._lib.rb_upload_medians(reg._ctx, np.ascontiguousarray(m).ctypes.data, 0, n))
            out = reg.filter_fragment(pl, mw, mh, want_fgmasks=True)
            want = oracle.filter_fragment(frames, m, pos, mw, mh)
            for i in range(n):
                assert np.array_equal(out["fgmasks"][i], want["masks"][i]), (case, which, W, H, colours, smooth, i)
            assert np.array_equal(out["ncontours"], want["ncontours"]), (case, which)
            assert np.array_equal(out["dots"], want["dots"]), (case, which)


@pytest.mark.parametrize("case", range(6))
def test_cell_match_random_maps(case):
    rng = np.random.default_rng(2000 + case)
    pw, ph = int(rng.integers(60, 500)), int(rng.integers(60, 400))
    cw, ch = int(rng.integers(60, 500)), int(rng.integers(60, 400))
    world = blobby(rng, 1, 900, 1100, int(rng.choice([3, 8, 16])), int(rng.choice([2, 4, 8])))[0]
    noise = rng.random(world.shape) < 0.04
    world = np.where(noise, rng.integers(0, 16, size=world.shape, dtype=np.uint8), world)

    def dots_of(img, holes):
        d = np.zeros(img.shape + (16,), np.uint16)
        np.put_along_axis(d, img[:, :, None].astype(np.int64), 1, axis=2)
        d[rng.random(img.shape) < holes] = 0          # never-visited pixels: mask 0
        return d

    ax, ay = int(rng.integers(0, 1100 - pw)), int(rng.integers(0, 900 - ph))
    bx, by = int(rng.integers(0, 1100 - cw)), int(rng.integers(0, 900 - ch))
    a, b = dots_of(world[ay:ay + ph, ax:ax + pw], 0.05), dots_of(world[by:by + ch, bx:bx + cw], 0.0)
    with remap_b200.Snippet(a) as sa, remap_b200.Snippet(b) as sb:
        oa, ob = oracle.snippet(a), oracle.snippet(b)
        for (s1, s2, o1, o2) in ((sa, sb, oa, ob), (sb, sa, ob, oa)):
            got, want = s1.match(s2), oracle.cell_match(o1, o2)
            for fld in ("offsets", "pairs", "ties", "matched_keypoints"):
                assert got[fld] == want[fld], (case, fld, got, want)
            if want["ties"] == 1:
                for fld in got.dtype.names:
                    assert got[fld] == want[fld], (case, fld, got, want)


@pytest.mark.parametrize("case", range(4))
def test_aws_compare_random(case):
    rng = np.random.default_rng(3000 + case)
    W, H, n = int(rng.integers(40, 420)), int(rng.integers(40, 330)), int(rng.integers(2, 120))
    base = rng.integers(0, 16, size=(H, W), dtype=np.uint8)
    frames = np.repeat(base[None], n, axis=0).copy()
    k = int(rng.integers(1, 400))
    frames[rng.integers(0, n, size=k), rng.integers(0, H, size=k), rng.integers(0, W, size=k)] = rng.integers(0, 256, size=k)
    try:
        reg = remap_b200.Registrar(W, H, max_frames=max(n, 2))
    except remap_b200.RemapError:
        pytest.skip("geometry")
    with reg:
        reg.upload(frames)
        heat, fc = reg.aws_compare(n)
    oh, ofc = oracle.aws_compare(frames)
    assert np.array_equal(heat, oh) and np.array_equal(fc, ofc)
