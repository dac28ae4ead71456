#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the REAL reference (oracle/_ref/ref_harness).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Each fixture holds the input frames and the reference's own canonical dump (oracle/ref_harness.cpp
`dump` mode: medians, per-region keypoints with 13-byte codes, per-region offset histograms and
tickets, declared offsets from the unmodified kpm::match, and positions from the unmodified
frc::collector loop).  The reference ships no fixtures of its own.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, refdump  # noqa: E402
from remap_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def raw_dump(frames):
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "f.bin"), os.path.join(td, "d.bin")
        frames.tofile(fin)
        subprocess.check_call([refdump.REF_BIN, "dump", fin, str(W), str(H), str(N), fout])
        return np.fromfile(fout, np.uint8)


def cases():
    yield "small_scroll", synth.scrolling_tilemap(8, 96, 64, seed=11, world_w=512, world_h=256, detail=2).frames
    yield "s_scroll", synth.scrolling_tilemap(4, 320, 224, seed=12).frames
    yield "random16", synth.random_frames(4, 96, 64, seed=13)
    yield "random3", synth.random_frames(4, 96, 64, seed=14, palette=3)
    yield "cuts", synth.scrolling_tilemap(7, 128, 96, seed=15, world_w=512, world_h=256, cut_every=3, levels=2).frames
    yield "odd", synth.scrolling_tilemap(5, 131, 99, seed=16, world_w=512, world_h=256).frames
    yield "flat", np.full((3, 96, 64), 5, np.uint8).reshape(3, 64, 96)
    yield "sprites", synth.scrolling_tilemap(5, 160, 112, seed=17, world_w=512, world_h=256, sprites=4).frames
    yield "parallax", synth.scrolling_tilemap(6, 160, 112, seed=18, world_w=512, world_h=256, parallax=16).frames
    # few tiles -> big groups of identical codes -> all-pairs blow-up (src/kpm.hpp:91-103)
    yield "repeat", synth.scrolling_tilemap(4, 160, 112, seed=19, world_w=512, world_h=256, n_tiles=3, speckle=0.02).frames


def raw_filter_dump(frames):
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "f.bin"), os.path.join(td, "d.bin")
        frames.tofile(fin)
        subprocess.check_call([refdump.REF_BIN, "filter", fin, str(W), str(H), str(N), fout, "1"],
                              stdout=subprocess.DEVNULL)
        return np.fromfile(fout, np.uint8)


def filter_cases():
    """Pass 2 (fdf::filter, src/fdf.hpp:40-91) fixtures: the reference's own collect + filter."""
    yield "filter_sprites", synth.scrolling_tilemap(8, 320, 224, seed=21, sprites=6, world_w=512, world_h=320).frames
    yield "filter_small", synth.scrolling_tilemap(16, 160, 112, seed=22, sprites=4, world_w=320, world_h=256).frames
    yield "filter_cuts", synth.scrolling_tilemap(14, 160, 112, seed=23, sprites=3, cut_every=5, levels=2,
                                                 world_w=320, world_h=256).frames
    rng = np.random.default_rng(24)   # a still camera and fresh noise every frame: many small contours
    base = synth.scrolling_tilemap(1, 160, 112, seed=24, world_w=320, world_h=256).frames[0]
    noisy = np.repeat(base[None], 10, axis=0).copy()
    for f in range(10):
        m = rng.random(base.shape) < 0.04
        noisy[f][m] = rng.integers(0, 16, size=int(m.sum()), dtype=np.uint8)
    yield "filter_noise", noisy


def raw_splice_dump(frames):
    N, H, W = frames.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "f.bin"), os.path.join(td, "d.bin")
        frames.tofile(fin)
        subprocess.check_call([refdump.REF_BIN, "splice", fin, str(W), str(H), str(N), fout], stdout=subprocess.DEVNULL)
        return np.fromfile(fout, np.uint8)


def splice_cases():
    """Fragment splicing (fgs::splice, src/fgs.hpp:187-213) fixtures: the reference's own collect, snippets,
    pairwise cellular kpm::match and splice."""
    yield "splice_levels", synth.scrolling_tilemap(60, 128, 96, seed=52, world_w=320, world_h=240, cut_every=15, levels=2).frames
    yield "splice_repeat", synth.scrolling_tilemap(90, 128, 96, seed=53, world_w=512, world_h=384, cut_every=18, n_tiles=4,
                                                   speckle=0.01).frames
    yield "splice_chain", synth.scrolling_tilemap(120, 160, 112, seed=41, world_w=400, world_h=304, cut_every=30).frames


def main():
    assert build_ref.build(), "needs /root/reference to build oracle/_ref"
    for name, frames in splice_cases():
        frames = np.ascontiguousarray(frames, np.uint8)
        dump = raw_splice_dump(frames)
        d = refdump.parse_splice_dump(dump.tobytes())
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), frames=frames, dump=dump)
        print(name, frames.shape, "fragments", len(d["fragments"]), "->", len(d["spliced"]), "matches",
              [(m["prev"], m["curr"], m["valid"], m["ties"]) for m in d["matches"]])
    if "--splice-only" in sys.argv:
        return
    # aws::details::compare (src/aws.hpp:37-60): a 96x64 "screen" with a static border around a scrolling window
    seq = synth.scrolling_tilemap(12, 64, 40, seed=25, world_w=256, world_h=128)
    screen = np.full((12, 64, 96), 6, np.uint8)
    screen[:, 10:50, 16:80] = seq.frames
    screen[5:, 2, 3] = 9           # a border pixel that changes once
    np.savez_compressed(os.path.join(OUT, "awsheat.npz"), frames=screen, heat=refdump.ref_heat(screen))
    print("awsheat ok")
    if "--heat-only" in sys.argv:
        return
    for name, frames in filter_cases():
        frames = np.ascontiguousarray(frames, np.uint8)
        dump = raw_filter_dump(frames)
        d = refdump.parse_filter_dump(dump.tobytes())
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), frames=frames, dump=dump)
        print(name, frames.shape, "fragments", len(d["fragments"]), "contours/frame",
              np.mean([len(f["contours"]) for f in d["frames"]]), "mask fraction", np.mean([f["mask"].mean() for f in d["frames"]]))
    if "--filter-only" in sys.argv:
        return
    for name, frames in cases():
        frames = np.ascontiguousarray(frames, np.uint8)
        dump = raw_dump(frames)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), frames=frames, dump=dump)
        print(name, frames.shape, "dump bytes", dump.size)
    # foreground mask fixtures (fde::details::generate_mask, src/fde.hpp:19-55)
    rng = np.random.default_rng(20)
    bg = rng.integers(0, 16, size=(150, 211), dtype=np.uint8)
    masks = {}
    for k, (px, py, W, H) in enumerate([(0, 0, 96, 64), (32, 5, 96, 64), (7, 3, 131, 99), (64, 0, 128, 96), (1, 1, 33, 7)]):
        frame = bg[py:py + H, px:px + W].copy()
        m = rng.random(frame.shape) < 0.2
        frame[m] = rng.integers(0, 16, size=int(m.sum()), dtype=np.uint8)
        masks[f"frame{k}"] = frame
        masks[f"pos{k}"] = np.array([px, py], np.int32)
        masks[f"mask{k}"] = refdump.ref_mask(bg, px, py, frame)
    np.savez_compressed(os.path.join(OUT, "fgmask.npz"), bg=bg, n=np.array(5), **masks)
    print("fgmask ok")


if __name__ == "__main__":
    main()
