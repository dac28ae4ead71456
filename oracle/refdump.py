"""Run oracle/_ref/ref_harness (the real reference, compiled) and parse its canonical dump.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (remap_b200/).
"""
from __future__ import annotations

import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "ref_harness")

KP_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("code", "u1", (13,))])
BIN_DTYPE = np.dtype([("dx", "<i4"), ("dy", "<i4"), ("cnt", "<u4")])
NREG = 8


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def ensure_ref():
    """Build oracle/_ref if the reference sources are present; return the binary or None."""
    from . import build_ref
    return build_ref.build(verbose=False)


def parse_dump(buf: bytes) -> dict:
    assert buf[:4] == b"RMDP"
    W, H, N = struct.unpack_from("<III", buf, 4)
    pos = 16
    frames = []
    pairs = []

    def read_grid(pos):
        regions = []
        for _ in range(NREG):
            n, w1, w2 = struct.unpack_from("<III", buf, pos)
            pos += 12
            kps = np.frombuffer(buf, KP_DTYPE, n, pos).copy()
            pos += n * KP_DTYPE.itemsize
            regions.append(dict(n=n, w1=w1, w2=w2, kps=kps))
        return regions, pos

    def read_pair(pos):
        valid, dx, dy, active = struct.unpack_from("<IiiI", buf, pos)
        pos += 16
        regs = []
        for _ in range(NREG):
            use_all, nb = struct.unpack_from("<II", buf, pos)
            pos += 8
            bins = np.frombuffer(buf, BIN_DTYPE, nb, pos).copy()
            pos += nb * BIN_DTYPE.itemsize
            (nt,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            ticket = np.frombuffer(buf, BIN_DTYPE, nt, pos).copy()
            pos += nt * BIN_DTYPE.itemsize
            regs.append(dict(use_all=bool(use_all), bins=bins, ticket=ticket))
        return dict(valid=bool(valid), dx=dx, dy=dy, active=active, regions=regs), pos

    for i in range(N):
        med = np.frombuffer(buf, np.uint8, W * H, pos).reshape(H, W).copy()
        pos += W * H
        regions, pos = read_grid(pos)
        frames.append(dict(median=med, regions=regions))
        if i > 0:
            p, pos = read_pair(pos)
            pairs.append(p)
    (nrec,) = struct.unpack_from("<I", buf, pos)
    pos += 4
    rec = np.frombuffer(buf, "<i4", nrec * 3, pos).reshape(nrec, 3).copy()
    pos += nrec * 12
    assert pos == len(buf), (pos, len(buf))
    return dict(W=W, H=H, N=N, frames=frames, pairs=pairs, positions=rec)


def ref_dump(frames: np.ndarray) -> dict:
    """frames: (N, H, W) uint8 -> parsed dump of the real reference's intermediates."""
    assert have_ref(), "oracle/_ref/ref_harness missing (run oracle/build_ref.py where /root/reference exists)"
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "frames.bin"), os.path.join(td, "dump.bin")
        np.ascontiguousarray(frames, np.uint8).tofile(fin)
        subprocess.check_call([REF_BIN, "dump", fin, str(W), str(H), str(N), fout])
        with open(fout, "rb") as f:
            return parse_dump(f.read())


def ref_bench(frames: np.ndarray, mode: str = "reg", threads: int = 1, reps: int = 1) -> dict:
    assert have_ref()
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin = os.path.join(td, "frames.bin")
        np.ascontiguousarray(frames, np.uint8).tofile(fin)
        out = subprocess.check_output([REF_BIN, "bench", fin, str(W), str(H), str(N), mode,
                                       str(threads), str(reps)])
    return json.loads(out.decode().strip().splitlines()[-1])


def ref_mask(bg: np.ndarray, px: int, py: int, frame: np.ndarray) -> np.ndarray:
    assert have_ref()
    bh, bw = bg.shape
    H, W = frame.shape
    with tempfile.TemporaryDirectory() as td:
        fb, ff, fo = (os.path.join(td, n) for n in ("bg.bin", "fr.bin", "mask.bin"))
        np.ascontiguousarray(bg, np.uint8).tofile(fb)
        np.ascontiguousarray(frame, np.uint8).tofile(ff)
        subprocess.check_call([REF_BIN, "mask", fb, str(bw), str(bh), str(px), str(py), ff, str(W), str(H), fo])
        return np.fromfile(fo, np.uint8).reshape(H, W)


CONTOUR_DTYPE = np.dtype([("area", "<u4"), ("left", "<u4"), ("top", "<u4"), ("right", "<u4"), ("bottom", "<u4"),
                          ("colour", "<u4")])


def parse_filter_dump(buf: bytes) -> dict:
    """Dump of `ref_harness filter`: the reference's frc::collector fragments run through fdf::filter."""
    assert buf[:4] == b"RMF2"
    W, H, N, nfrag = struct.unpack_from("<IIII", buf, 4)
    pos = 20
    backgrounds = []
    for _ in range(nfrag):
        bw, bh, zx, zy = struct.unpack_from("<IIii", buf, pos)
        pos += 16
        img = np.frombuffer(buf, np.uint8, bw * bh, pos).reshape(bh, bw).copy()
        pos += bw * bh
        backgrounds.append(dict(image=img, zero=(zx, zy)))
    frames = []
    for _ in range(N):
        frag, no, x, y, nc = struct.unpack_from("<IIiiI", buf, pos)
        pos += 20
        cont = np.frombuffer(buf, CONTOUR_DTYPE, nc, pos).copy()
        pos += nc * CONTOUR_DTYPE.itemsize
        mask = np.frombuffer(buf, np.uint8, W * H, pos).reshape(H, W).copy()
        pos += W * H
        frames.append(dict(fragment=frag, number=no, x=x, y=y, contours=cont, mask=mask))
    fragments = []
    for _ in range(nfrag):
        bw, bh, zx, zy = struct.unpack_from("<IIii", buf, pos)
        pos += 16
        dots = np.frombuffer(buf, "<u2", bw * bh * 16, pos).reshape(bh, bw, 16).copy()
        pos += bw * bh * 32
        fragments.append(dict(dots=dots, zero=(zx, zy)))
    assert pos == len(buf), (pos, len(buf))
    return dict(W=W, H=H, N=N, backgrounds=backgrounds, frames=frames, fragments=fragments)


def ref_filter(frames: np.ndarray, reps: int = 1) -> dict:
    """frames (N, H, W) -> the real reference's collect + fdf::filter results (+ 'timing')."""
    assert have_ref()
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "frames.bin"), os.path.join(td, "dump.bin")
        np.ascontiguousarray(frames, np.uint8).tofile(fin)
        out = subprocess.check_output([REF_BIN, "filter", fin, str(W), str(H), str(N), fout, str(reps)])
        with open(fout, "rb") as f:
            d = parse_filter_dump(f.read())
    d["timing"] = json.loads(out.decode().strip().splitlines()[-1])
    return d


def _read_fragment(buf, pos):
    w, h, zx, zy, nf = struct.unpack_from("<IIiiI", buf, pos)
    pos += 20
    fr = np.frombuffer(buf, "<i4", nf * 3, pos).reshape(nf, 3).copy()   # number, x, y
    pos += nf * 12
    dots = np.frombuffer(buf, "<u2", w * h * 16, pos).reshape(h, w, 16).copy()
    pos += w * h * 32
    return dict(zero=(zx, zy), frames=fr, dots=dots), pos


SNIP_KP_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("code", "u1", (13,))])


def parse_splice_dump(buf: bytes) -> dict:
    """Dump of `ref_harness splice`: fragments, fgs snippets, every pairwise cellular kpm::match, fgs::splice."""
    assert buf[:4] == b"RMSP"
    W, H, N, nfrag = struct.unpack_from("<IIII", buf, 4)
    pos = 20
    fragments = []
    for _ in range(nfrag):
        f, pos = _read_fragment(buf, pos)
        fragments.append(f)
    snippets = []
    for _ in range(nfrag):
        w, h, nk = struct.unpack_from("<III", buf, pos)
        pos += 12
        kps = np.frombuffer(buf, SNIP_KP_DTYPE, nk, pos).copy()
        pos += nk * SNIP_KP_DTYPE.itemsize
        mask = np.frombuffer(buf, np.uint8, w * h, pos).reshape(h, w).copy()
        pos += w * h
        snippets.append(dict(kps=kps, mask=mask))
    matches = []
    for i in range(nfrag):
        for j in range(i + 1, nfrag):
            a, b, noff = struct.unpack_from("<III", buf, pos)
            pos += 12
            (pairs,) = struct.unpack_from("<Q", buf, pos)
            pos += 8
            (ties,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            m = dict(prev=a, curr=b, offsets=noff, pairs=pairs, ties=ties)
            if noff:
                dx, dy, kp, cells, active = struct.unpack_from("<iiIII", buf, pos)
                pos += 20
                m.update(best=(dx, dy), matched_keypoints=kp, matched_cells=cells, active_cells=active)
            valid, vx, vy, cnt = struct.unpack_from("<IiiI", buf, pos)
            pos += 16
            m.update(valid=bool(valid), vote=(vx, vy), count=cnt)
            matches.append(m)
    (ns,) = struct.unpack_from("<I", buf, pos)
    pos += 4
    spliced = []
    for _ in range(ns):
        f, pos = _read_fragment(buf, pos)
        spliced.append(f)
    assert pos == len(buf), (pos, len(buf))
    return dict(W=W, H=H, N=N, fragments=fragments, snippets=snippets, matches=matches, spliced=spliced)


def ref_splice(frames: np.ndarray) -> dict:
    assert have_ref()
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "frames.bin"), os.path.join(td, "dump.bin")
        np.ascontiguousarray(frames, np.uint8).tofile(fin)
        out = subprocess.check_output([REF_BIN, "splice", fin, str(W), str(H), str(N), fout])
        with open(fout, "rb") as f:
            d = parse_splice_dump(f.read())
    d["timing"] = json.loads(out.decode().strip().splitlines()[-1])
    return d


def ref_heat(frames: np.ndarray) -> np.ndarray:
    """aws::details::compare over consecutive pairs -> (N-1, H, W) heat maps after each pair (W*H % 32 == 0)."""
    assert have_ref()
    N, H, W = frames.shape
    assert (W * H) % 32 == 0, "compare's vector loop runs to the next 32-byte boundary (src/aws.hpp:45-52)"
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "frames.bin"), os.path.join(td, "heat.bin")
        np.ascontiguousarray(frames, np.uint8).tofile(fin)
        subprocess.check_call([REF_BIN, "heat", fin, str(W), str(H), str(N), fout])
        return np.fromfile(fout, np.uint8).reshape(N - 1, H, W)
