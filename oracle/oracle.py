"""ctypes wrapper of oracle/remap_oracle.c (the CPU restatement).  TEST INFRASTRUCTURE ONLY.

Never imported by the product path (remap_b200/); only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg use it, as the checker.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "remap_oracle.c")
LIB = os.path.join(HERE, "_build", "libremap_oracle.so")

KP_DTYPE = np.dtype([("code", "u1", (13,)), ("weight", "u1"), ("x", "<u2"), ("y", "<u2"),
                     ("region_mask", "<u4")], align=True)
BIN_DTYPE = np.dtype([("dx", "<i4"), ("dy", "<i4"), ("cnt", "<u4")])
VOTE_DTYPE = np.dtype([("use_all", "<u4"), ("n_prev", "<u4"), ("n_curr", "<u4"), ("w2_prev", "<u4"),
                       ("w2_curr", "<u4"), ("nbins", "<u4"), ("nticket", "<u4"),
                       ("ticket", BIN_DTYPE, (4,)), ("ngt", "<u4", (4,)), ("nge", "<u4", (4,)), ("hist_hash", "<u4")])
RESULT_DTYPE = np.dtype([("dx", "<i4"), ("dy", "<i4"), ("valid", "<u4"), ("tie_sensitive", "<u4"),
                         ("active", "<u4"), ("top_dx", "<i4", (2,)), ("top_dy", "<i4", (2,)),
                         ("top_score", "<u4", (2,)), ("ntop", "<u4")])
CONTOUR_DTYPE = np.dtype([("area", "<u4"), ("left", "<u4"), ("top", "<u4"), ("right", "<u4"), ("bottom", "<u4"),
                          ("colour", "<u4")])
CELL_MATCH_DTYPE = np.dtype([("valid", "<u4"), ("dx", "<i4"), ("dy", "<i4"), ("matched_keypoints", "<u4"),
                             ("matched_cells", "<u4"), ("active_cells", "<u4"), ("offsets", "<u4"), ("ties", "<u4"),
                             ("pairs", "<u8")])


class Config(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("width", "height", "grid_w", "grid_h", "overlap", "weight_switch", "region_votes")]


def config(width, height, grid_w=4, grid_h=2, overlap=16, weight_switch=10, region_votes=3):
    """Defaults are the reference's constants: src/frc.hpp:22-24,32-33."""
    return Config(width, height, grid_w, grid_h, overlap, weight_switch, region_votes)


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ro_extract.restype = C.c_size_t
        _lib.ro_extract.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.ro_match.restype = None
        _lib.ro_match.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                  C.c_void_p, C.c_void_p]
        _lib.ro_region_bins.restype = C.c_size_t
        _lib.ro_region_bins.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_uint32, C.c_void_p, C.c_size_t]
        _lib.ro_register.restype = C.c_size_t
        _lib.ro_register.argtypes = [C.POINTER(Config), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
        _lib.ro_foreground_mask.restype = None
        _lib.ro_foreground_mask.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.ro_foreground.restype = C.c_size_t
        _lib.ro_foreground.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.ro_cell_match.restype = None
        _lib.ro_cell_match.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.ro_luts.restype = None
        _lib.ro_luts.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ro_sections.restype = None
        _lib.ro_sections.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p]
        for fn, dt in (("ro_sizeof_keypoint", KP_DTYPE), ("ro_sizeof_region_vote", VOTE_DTYPE),
                       ("ro_sizeof_match_result", RESULT_DTYPE), ("ro_sizeof_contour", CONTOUR_DTYPE),
                       ("ro_sizeof_cell_match", CELL_MATCH_DTYPE)):
            f = getattr(_lib, fn)
            f.restype = C.c_size_t
            assert f() == dt.itemsize, (fn, f(), dt.itemsize)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def luts():
    n2o = np.zeros(16, np.uint8)
    o2n = np.zeros(16, np.uint8)
    lib().ro_luts(_p(n2o), _p(o2n))
    return n2o, o2n


def sections(cfg):
    cs = np.zeros(cfg.width, np.uint32)
    rs = np.zeros(cfg.height, np.uint32)
    lib().ro_sections(C.byref(cfg), _p(cs), _p(rs))
    return cs, rs


def extract(cfg, frame):
    """-> (median (H, W) u8, keypoints structured array in the reference's insertion order)."""
    frame = np.ascontiguousarray(frame, np.uint8)
    H, W = frame.shape
    assert (W, H) == (cfg.width, cfg.height)
    median = np.zeros((H, W), np.uint8)
    kps = np.zeros(W * H, KP_DTYPE)
    n = lib().ro_extract(C.byref(cfg), _p(frame), _p(median), _p(kps), kps.shape[0])
    return median, kps[:n].copy()


def match(cfg, prev_kps, curr_kps):
    """-> (result record, per-region vote records)."""
    prev_kps = np.ascontiguousarray(prev_kps)
    curr_kps = np.ascontiguousarray(curr_kps)
    res = np.zeros(1, RESULT_DTYPE)
    votes = np.zeros(cfg.grid_w * cfg.grid_h, VOTE_DTYPE)
    lib().ro_match(C.byref(cfg), _p(prev_kps), prev_kps.shape[0], _p(curr_kps), curr_kps.shape[0],
                   _p(res), _p(votes))
    return res[0], votes


def region_bins(cfg, prev_kps, curr_kps, region):
    prev_kps = np.ascontiguousarray(prev_kps)
    curr_kps = np.ascontiguousarray(curr_kps)
    cap = 1 << 16
    while True:
        bins = np.zeros(cap, BIN_DTYPE)
        n = lib().ro_region_bins(C.byref(cfg), _p(prev_kps), prev_kps.shape[0], _p(curr_kps),
                                 curr_kps.shape[0], region, _p(bins), cap)
        if n <= cap:
            return bins[:n].copy()
        cap = int(n)


def register(cfg, frames, want_medians=False):
    """The frc loop over a whole sequence.
    -> dict(results (N-1,) RESULT_DTYPE, positions (N,3) i32 [fragment,x,y], medians|None, kp_counts (N,))"""
    frames = np.ascontiguousarray(frames, np.uint8)
    N, H, W = frames.shape
    assert (W, H) == (cfg.width, cfg.height)
    results = np.zeros(max(N - 1, 0), RESULT_DTYPE)
    positions = np.zeros((N, 3), np.int32)
    medians = np.zeros((N, H, W), np.uint8) if want_medians else None
    kpc = np.zeros(N, np.uint32)
    total = lib().ro_register(C.byref(cfg), _p(frames), N, _p(results), _p(positions), _p(medians), _p(kpc))
    return dict(results=results, positions=positions, medians=medians, kp_counts=kpc, total_keypoints=int(total))


def foreground_mask(bg, px, py, frame):
    bg = np.ascontiguousarray(bg, np.uint8)
    frame = np.ascontiguousarray(frame, np.uint8)
    bh, bw = bg.shape
    H, W = frame.shape
    mask = np.zeros((H, W), np.uint8)
    lib().ro_foreground_mask(_p(bg), bw, bh, px, py, _p(frame), W, H, _p(mask))
    return mask


def assemble_fragment(frames, pos):
    """TEST ORACLE for map assembly: fgm::fragment::blit of every frame, one by one, with the map growing
    exactly as the reference grows it (src/fgm.hpp:87-97,176-233; mrl::matrix::extend src/mrl.hpp:131-147),
    then fragment::blend (src/fgm.hpp:115-135).  Plain numpy, small cases only.
    frames: (n, H, W) uint8 0..15, pos: (n, 2) positions of one fragment.
    -> dict(zero=(x, y), dots (mh, mw, 16) uint16, image, mask)"""
    n, H, W = frames.shape
    zero = [0, 0]
    dots = np.zeros((H, W, 16), np.uint16)
    ar = np.arange(16, dtype=np.uint8)
    for f in range(n):
        px, py = int(pos[f][0]), int(pos[f][1])
        grow = [0, 0, 0, 0]  # left, top, right, bottom
        for k, (p, step) in enumerate(((px, W), (py, H))):
            dim = dots.shape[1 - k]
            if p < zero[k]:
                ch = zero[k] - p
                grow[k] = ch - ch % step + (step if ch % step else 0)
            required = p + step
            if required > 0 and required > zero[k] + dim:
                ch = required - (zero[k] + dim)
                grow[k + 2] = ch - ch % step + (step if ch % step else 0)
            zero[k] -= grow[k]
        if any(grow):
            nd = np.zeros((dots.shape[0] + grow[1] + grow[3], dots.shape[1] + grow[0] + grow[2], 16), np.uint16)
            nd[grow[1]:grow[1] + dots.shape[0], grow[0]:grow[0] + dots.shape[1]] = dots
            dots = nd
        ax, ay = px - zero[0], py - zero[1]
        onehot = (frames[f][:, :, None] == ar[None, None, :]).astype(np.uint16)
        dots[ay:ay + H, ax:ax + W] += onehot  # uint16 arithmetic wraps like the reference's counters
    best = dots.max(axis=2)
    image = np.where(best != 0, dots.argmax(axis=2), 0).astype(np.uint8)  # argmax = first largest
    mask = (best != 0).astype(np.uint8)
    return dict(zero=(zero[0], zero[1]), dots=dots, image=image, mask=mask)


def foreground(bg, px, py, frame, median):
    """Pass-2 foreground of one frame: fde::extractor::extract + fde::mask (src/fde.hpp:83-146).
    -> (mask (H, W) u8 0/1, kept contours in the reference's order)"""
    bg = np.ascontiguousarray(bg, np.uint8)
    frame = np.ascontiguousarray(frame, np.uint8)
    median = np.ascontiguousarray(median, np.uint8)
    bh, bw = bg.shape
    H, W = frame.shape
    assert 0 <= px and px + W <= bw and 0 <= py and py + H <= bh
    mask = np.zeros((H, W), np.uint8)
    cont = np.zeros(W * H, CONTOUR_DTYPE)
    n = lib().ro_foreground(_p(bg), bw, bh, px, py, _p(frame), _p(median), W, H, _p(mask), _p(cont), cont.shape[0])
    assert n != 2 ** 64 - 1, "more than 65,534 contours in one frame: undefined in the reference"
    return mask, cont[:n].copy()


def filter_fragment(frames, medians, pos, mapW, mapH, background=None):
    """TEST ORACLE for fdf::filter over ONE fragment (src/fdf.hpp:40-75): background = blend of the plain
    blit of all frames (src/fdf.hpp:21-34) unless given; per frame the pass-2 foreground mask, then the
    masked blit (src/fgm.hpp:71-85: ++dots[colour] where mask == 0).  pos: (n, 2) positions inside the map
    (frame position minus fragment zero).  numpy + the C restatement; small cases only.
    -> dict(background, masks (n, H, W), ncontours (n,), dots (mapH, mapW, 16) u16)"""
    n, H, W = frames.shape
    ar = np.arange(16, dtype=np.uint8)
    if background is None:
        d0 = np.zeros((mapH, mapW, 16), np.uint16)
        for f in range(n):
            x, y = int(pos[f][0]), int(pos[f][1])
            d0[y:y + H, x:x + W] += (frames[f][:, :, None] == ar).astype(np.uint16)
        best = d0.max(axis=2)
        background = np.where(best != 0, d0.argmax(axis=2), 0).astype(np.uint8)
    dots = np.zeros((mapH, mapW, 16), np.uint16)
    masks = np.zeros((n, H, W), np.uint8)
    nc = np.zeros(n, np.uint32)
    for f in range(n):
        x, y = int(pos[f][0]), int(pos[f][1])
        m, cont = foreground(background, x, y, frames[f], medians[f])
        masks[f] = m
        nc[f] = len(cont)
        dots[y:y + H, x:x + W] += ((frames[f][:, :, None] == ar) & (m[:, :, None] == 0)).astype(np.uint16)
    return dict(background=background, masks=masks, ncontours=nc, dots=dots)


def blend(dots):
    """fgm::fragment::blend (src/fgm.hpp:115-135): -> (image, mask)"""
    best = dots.max(axis=2)
    return np.where(best != 0, dots.argmax(axis=2), 0).astype(np.uint8), (best != 0).astype(np.uint8)


def snippet(dots):
    """fgs::details::extract_single (src/fgs.hpp:80-89): blend + kpe with a 1 x 1 grid, no overlap.
    -> dict(image, mask, kps)"""
    image, mask = blend(dots)
    H, W = image.shape
    _, kps = extract(config(W, H, 1, 1, 0), image)
    return dict(image=image, mask=mask, kps=kps)


def cell_match(prev, curr, cell=(15, 15)):
    """The cellular kpm::match (src/kpm.hpp:371-393) of two snippets (dicts from snippet())."""
    pk, ck = np.ascontiguousarray(prev["kps"]), np.ascontiguousarray(curr["kps"])
    pm = np.ascontiguousarray(prev["mask"], np.uint8)
    res = np.zeros(1, CELL_MATCH_DTYPE)
    lib().ro_cell_match(_p(pk), len(pk), _p(pm), pm.shape[1], pm.shape[0], _p(ck), len(ck), curr["mask"].shape[1],
                        curr["mask"].shape[0], cell[0], cell[1], _p(res))
    return res[0]


def aws_compare(frames, heat=None):
    """aws::details::compare (src/aws.hpp:37-60) applied to every consecutive pair: heat &= (prev == curr).
    -> (heat after all pairs, first_change (H, W) uint32: index of the first differing pair, 0xFFFFFFFF if none)"""
    frames = np.asarray(frames, np.uint8)
    N, H, W = frames.shape
    h = np.ones((H, W), np.uint8) if heat is None else np.array(heat, np.uint8)
    fc = np.full((H, W), 0xFFFFFFFF, np.uint32)
    for i in range(N - 1):
        ne = frames[i] != frames[i + 1]
        fc[ne & (fc == 0xFFFFFFFF)] = i
        h[ne] = 0
    return h, fc
