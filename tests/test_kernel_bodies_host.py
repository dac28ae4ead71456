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
