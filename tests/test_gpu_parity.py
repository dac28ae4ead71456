"""-m gpu: the CUDA path (through the C ABI) against (a) committed dumps of the REAL reference,
(b) the C restatement on seeded inputs, (c) size-independent properties at BASELINE sizes."""
import glob
import os

import numpy as np
import pytest

import gpu_common
import parity
import remap_b200
from oracle import oracle, refdump
from remap_b200 import RB_OFFSET_TIE_SENSITIVE, RB_OFFSET_VALID, synth

pytestmark = pytest.mark.gpu

GOLDEN = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not p.endswith(("fgmask.npz", "awsheat.npz")) and not os.path.basename(p).startswith(("filter_", "splice_")))


@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_matches_reference_dump(name, golden_dir):
    """Every intermediate the reference produced for this fixture, byte for byte."""
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    frames = z["frames"]
    ref = refdump.parse_dump(z["dump"].tobytes())
    n = frames.shape[0]
    out = gpu_common.run_sequence(frames)
    reg = out["reg"]
    try:
        valid, dx, dy, flagged = [], [], [], 0
        for i in range(n):
            parity.check_frame(ref["frames"][i], out["medians"][i], reg.keypoints(i), f"{name} frame {i}")
            if i > 0:
                ballots = reg.region_ballots(i - 1)
                bins = [reg.region_votes(i - 1, r) for r in range(8)]
                rec = gpu_common.result_record(out["offsets"][i - 1], ballots)
                st = parity.check_pair(ref["pairs"][i - 1], rec, ballots, bins, f"{name} pair {i}")
                flagged += st == "flagged"
                valid.append(rec["valid"]); dx.append(rec["dx"]); dy.append(rec["dy"])
        if flagged == 0:  # positions of the unmodified frc::collector loop
            assert np.array_equal(parity.positions_from_results(valid, dx, dy), ref["positions"])
        if name != "parallax":
            assert flagged == 0
    finally:
        reg.close()


ORACLE_CASES = {
    "S": (lambda: synth.scrolling_tilemap(6, 320, 224, seed=31).frames, {}),
    "S_dense": (lambda: synth.scrolling_tilemap(4, 320, 224, seed=32, speckle=0.10, detail=3).frames, {}),
    "S_repeat": (lambda: synth.scrolling_tilemap(4, 320, 224, seed=33, speckle=0.10, n_tiles=4).frames, {}),
    "S_small_tables": (lambda: synth.scrolling_tilemap(4, 320, 224, seed=34, speckle=0.10, n_tiles=4).frames,
                       dict(code_slots=256, offset_slots=1024)),
    "L": (lambda: synth.scrolling_tilemap(3, 640, 480, seed=35, speckle=0.10, vmax=(48, 48)).frames, {}),
    "odd": (lambda: synth.scrolling_tilemap(4, 323, 227, seed=36).frames, {}),
    "random16": (lambda: synth.random_frames(3, 320, 224, seed=37), {}),
    "random2": (lambda: synth.random_frames(3, 160, 112, seed=38, palette=2), {}),
    "cuts": (lambda: synth.scrolling_tilemap(8, 320, 224, seed=39, cut_every=3, levels=2).frames, {}),
    "sprites": (lambda: synth.scrolling_tilemap(5, 320, 224, seed=40, sprites=12).frames, {}),
    "zeros": (lambda: np.zeros((3, 224, 320), np.uint8), {}),
    # tall regions: the matcher's tile needs two stacked TMA boxes (more than 256 rows per region)
    "tall_1024x768": (lambda: synth.scrolling_tilemap(3, 1024, 768, seed=43, speckle=0.01, vmax=(20, 15)).frames, {}),
    "tall_320x600": (lambda: synth.scrolling_tilemap(4, 320, 600, seed=44, world_h=1024).frames, {}),
    "one_region_active": (lambda: _one_corner_textured(), {}),
    "two_frames": (lambda: synth.scrolling_tilemap(2, 320, 224, seed=41).frames, {}),
}


def _one_corner_textured():
    """texture only in the top-left corner: a single active region -> kpm::match gives up (src/kpm.hpp:401)"""
    f = synth.scrolling_tilemap(3, 320, 224, seed=45).frames.copy()
    f[:, :, 80:] = 0
    f[:, 110:, :] = 0
    return f


# matcher variants: the pipelined matcher (default), the general kernel alone, the pipelined matcher with
# a list capacity so small that most regions are deferred to the general kernel, and short runs
MATCHER_MODES = {
    "pipelined": {},
    "general": dict(kpm_mode=1),
    "tiny_cap": dict(list_cap=64),
    "run3": dict(run_pairs=3),
    "overlap3": dict(overlap_batches=3),
    "big_first": dict(kpm_mode=2),            # the large-region kernel (rb_kpm_big.cuh) as the first pass
    "big_second": dict(list_cap=128),         # ... and as the second pass behind a small first-pass capacity
}


@pytest.mark.parametrize("mode", sorted(MATCHER_MODES))
@pytest.mark.parametrize("name", sorted(ORACLE_CASES))
def test_cuda_matches_oracle(name, mode):
    make, kw = ORACLE_CASES[name]
    frames = make()
    kw = dict(kw, **MATCHER_MODES[mode])
    gpu_common.compare_with_oracle(frames, oracle, f"{name}/{mode}", taps=((0, 0), (0, 3), (0, 7), (frames.shape[0] - 2, 5)),
                                   **kw)


def test_pipelined_matcher_equals_general_kernel_ballot_for_ballot():
    """Same frames through both matchers: every field of every region ballot of every pair."""
    n = 200
    seq = synth.scrolling_tilemap(n, 320, 224, seed=77, cut_every=40, sprites=6)
    outs = []
    for kw in (dict(), dict(kpm_mode=1), dict(list_cap=256, run_pairs=7)):
        with remap_b200.Registrar(320, 224, max_frames=n, **kw) as reg:
            reg.upload(seq.frames)
            off, _ = reg.register(n)
            ballots = np.stack([reg.region_ballots(i) for i in range(n - 1)])
            outs.append((off.copy(), ballots, reg.deferred_count))
    # scene cuts make thousands of spurious offsets in a region: those few pairs overflow the offset table
    assert outs[0][2] < 40, "the pipelined matcher should defer next to nothing on a scrolling workload"
    assert outs[2][2] > 0, "list_cap=256 should exercise the deferral path"
    for off, ballots, _ in outs[1:]:
        assert np.array_equal(off, outs[0][0])
        for fld in ballots.dtype.names:
            assert np.array_equal(ballots[fld], outs[0][1][fld]), fld


@pytest.mark.parametrize("size,speckle,vmax", [((640, 480), 0.10, (48, 48)), ((640, 480), 0.30, (20, 20)), ((512, 384), 0.10, (30, 30))])
def test_large_region_matcher_equals_general_kernel(size, speckle, vmax):
    """640x480-class frames: 4-5 k keypoints per region go through rb_kpm_big_kernel; every field of every ballot
    (histogram digest included) and every declared offset must equal the general kernel's, and nothing may be
    deferred at the density of BASELINE configs[3]."""
    n = 48
    W, H = size
    seq = synth.scrolling_tilemap(n, W, H, seed=404, speckle=speckle, vmax=vmax, cut_every=19)
    outs = []
    for kw in (dict(), dict(kpm_mode=1), dict(run_pairs=5)):
        with remap_b200.Registrar(W, H, max_frames=n, **kw) as reg:
            reg.upload(seq.frames)
            off, _ = reg.register(n)
            ballots = np.stack([reg.region_ballots(i) for i in range(n - 1)])
            outs.append((off.copy(), ballots, reg.deferred_count))
    if speckle <= 0.10:
        assert outs[0][2] <= 16, f"large-region matcher deferred {outs[0][2]} ballots"
    assert int(outs[0][1]["n_curr"].max()) > 2047, "the case should exceed the first-pass kernel's lists"
    for off, ballots, _ in outs[1:]:
        assert np.array_equal(off, outs[0][0])
        for fld in ballots.dtype.names:
            assert np.array_equal(ballots[fld], outs[0][1][fld]), fld


def test_large_region_matcher_epoch_wrap(monkeypatch):
    """The bucket heads of rb_kpm_big_kernel carry a 16-bit epoch instead of being cleared; started just below the
    wrap, every CTA crosses it within a few work items and must zero its tables exactly once."""
    n = 120
    seq = synth.scrolling_tilemap(n, 640, 480, seed=405, speckle=0.08, vmax=(30, 30))
    with remap_b200.Registrar(640, 480, max_frames=n, run_pairs=4) as reg:
        reg.upload(seq.frames)
        off_a, _ = reg.register(n)
        ballots_a = reg.fetch_ballots(n - 1)
    monkeypatch.setenv("RB_BIG_EPOCH0", "65530")  # first work item fits below the wrap, the second crosses it
    with remap_b200.Registrar(640, 480, max_frames=n, run_pairs=4) as reg:
        assert reg.matcher_kernel == "rb_kpm_big_kernel"
        reg.upload(seq.frames)
        off_b, _ = reg.register(n)
        ballots_b = reg.fetch_ballots(n - 1)
        assert reg.deferred_count == 0
    assert np.array_equal(off_a, off_b)
    for fld in ballots_a.dtype.names:
        assert np.array_equal(ballots_a[fld], ballots_b[fld]), fld
    assert np.array_equal(np.stack([off_a["dx"], off_a["dy"]], 1), seq.true_offsets)


def test_pipelined_matcher_defers_nothing_on_config2_frames():
    n = 500
    seq = synth.scrolling_tilemap(n, 320, 224, seed=1)
    with remap_b200.Registrar(320, 224, max_frames=n) as reg:
        reg.upload(seq.frames)
        reg.register(n)
        assert reg.deferred_count == 0


@pytest.mark.parametrize("chunk,size", [(3, (320, 224)), (64, (320, 224)), (0, (320, 224)), (17, (323, 227)), (40, (200, 136))])
def test_pipelined_host_registration_equals_upload_then_register(chunk, size):
    """rb_register_host_async (chunked copies overlapped with the kernels, one-frame overlap between
    chunks) gives the same offsets, ballots and medians as rb_upload + rb_register."""
    n = 150
    W, H = size
    seq = synth.scrolling_tilemap(n, W, H, seed=91, cut_every=35)
    with remap_b200.Registrar(W, H, max_frames=n + 5) as reg:
        reg.upload(seq.frames, first=2)
        off_a, med_a = reg.register(n, first=2, want_medians=True)
        ballots_a = np.stack([reg.region_ballots(i) for i in range(n - 1)])
    with remap_b200.Registrar(W, H, max_frames=n + 5, upload_chunk=chunk) as reg:
        reg.register_host_async(seq.frames | 0x50, first=2)  # dirty high nibbles must not matter
        off_b = reg.fetch_offsets(n - 1)
        med_b = reg.fetch_medians(n, first=2)
        ballots_b = np.stack([reg.region_ballots(i) for i in range(n - 1)])
    assert np.array_equal(off_a, off_b) and np.array_equal(med_a, med_b)
    for fld in ballots_a.dtype.names:
        assert np.array_equal(ballots_a[fld], ballots_b[fld]), fld


@pytest.mark.parametrize("lane", ["raw", "packed", "auto"])
def test_host_registration_lanes_from_pinned_memory(lane, monkeypatch):
    """Page-locked host frames may travel raw (one copy, packed on the device) or packed by host threads, chunk by
    chunk (rb_register_host_async picks from measured rates; RB_HOST_LANE forces one): same results either way."""
    import torch
    n, W, H = 300, 320, 224
    seq = synth.scrolling_tilemap(n, W, H, seed=92, cut_every=70)
    pinned = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[...] = seq.frames
    with remap_b200.Registrar(W, H, max_frames=n) as reg:
        reg.upload(seq.frames)
        off_a, med_a = reg.register(n, want_medians=True)
    if lane != "auto":
        monkeypatch.setenv("RB_HOST_LANE", lane)
    with remap_b200.Registrar(W, H, max_frames=n, upload_chunk=32) as reg:
        for _ in range(2):  # the second call runs on the rates measured by the first
            reg.register_host_async(pinned.numpy())
            off_b = reg.fetch_offsets(n - 1)
        med_b = reg.fetch_medians(n)
        st = reg.host_lane_stats
    assert np.array_equal(off_a, off_b) and np.array_equal(med_a, med_b)
    assert st["raw_chunks"] + st["packed_chunks"] == (n + 31) // 32
    if lane == "raw":
        assert st["packed_chunks"] == 0
    if lane == "packed":
        assert st["raw_chunks"] == 0


@pytest.mark.parametrize("size,pad", [((320, 224), 0), ((323, 227), 0), ((200, 136), 12)])
def test_registration_of_caller_packed_frames(size, pad):
    """rb_register_host_packed4: the caller already holds 4 bit/pixel rows (any row pitch >= ceil(W / 2))."""
    n = 90
    W, H = size
    seq = synth.scrolling_tilemap(n, W, H, seed=93)
    rb = (W + 1) // 2 + pad
    lo = np.zeros((n, H, 2 * rb), np.uint8)
    lo[:, :, :W] = seq.frames
    packed = (lo[:, :, 0::2] | (lo[:, :, 1::2] << 4)).astype(np.uint8)
    if pad:
        packed[:, :, (W + 1) // 2:] = 0xEE  # bytes beyond the row are the caller's and must be ignored
    with remap_b200.Registrar(W, H, max_frames=n) as reg:
        reg.upload(seq.frames)
        off_a, med_a = reg.register(n, want_medians=True)
    with remap_b200.Registrar(W, H, max_frames=n, upload_chunk=25) as reg:
        reg.register_host_packed4(packed)
        off_b = reg.fetch_offsets(n - 1)
        med_b = reg.fetch_medians(n)
    assert np.array_equal(off_a, off_b) and np.array_equal(med_a, med_b)


def test_single_frame_and_two_frame_registrations():
    """n = 1: extraction only, no pair; n = 2: one pair; through both entry points."""
    seq = synth.scrolling_tilemap(2, 320, 224, seed=46)
    cfg = oracle.config(320, 224)
    med0, kps0 = oracle.extract(cfg, seq.frames[0])
    with remap_b200.Registrar(320, 224, max_frames=4) as reg:
        reg.upload(seq.frames[:1])
        off, med = reg.register(1, want_medians=True)
        assert len(off) == 0 and np.array_equal(med[0], med0)
        assert np.array_equal(reg.keypoints(0)["code"], kps0["code"])
        reg.register_host_async(seq.frames[:1], first=1)
        assert len(reg.fetch_offsets(0)) == 0
        assert np.array_equal(reg.fetch_medians(1, first=1)[0], med0)
        reg.register_host_async(seq.frames, first=2)
        off2 = reg.fetch_offsets(1)
        assert (int(off2["dx"][0]), int(off2["dy"][0])) == tuple(int(v) for v in seq.true_offsets[0])
        assert off2["flags"][0] & RB_OFFSET_VALID


def test_dirty_high_nibbles_are_ignored():
    """The reference requires 0..15 (src/cpl.hpp LUT index); the kernels mask to the low nibble."""
    frames = synth.scrolling_tilemap(3, 320, 224, seed=42).frames
    a = gpu_common.run_sequence(frames)
    b = gpu_common.run_sequence(frames | 0xA0)
    try:
        assert np.array_equal(a["offsets"], b["offsets"]) and np.array_equal(a["medians"], b["medians"])
    finally:
        a["reg"].close(); b["reg"].close()


def test_full_size_config2_properties():
    """BASELINE config 2: 320x224, 20,000 frames on one GPU.  Size-independent checks:
    ground-truth camera deltas, range-split invariance (== the multi-GPU sharding rule), determinism,
    and a sampled diff against the oracle."""
    n = 20000
    seq = synth.scrolling_tilemap(n, 320, 224, seed=1)
    with remap_b200.Registrar(320, 224, max_frames=n) as reg:
        reg.upload(seq.frames)
        off, _ = reg.register(n)
        assert (off["flags"] & RB_OFFSET_VALID).all()
        assert not (off["flags"] & RB_OFFSET_TIE_SENSITIVE).any()
        assert np.array_equal(np.stack([off["dx"], off["dy"]], 1), seq.true_offsets)
        # determinism
        off2, _ = reg.register(n)
        assert np.array_equal(off, off2)
        # contiguous ranges with a one-frame overlap give the same pairs
        cut = 7777
        a, _ = reg.register(cut + 1, first=0)
        a = a.copy()
        b, _ = reg.register(n - cut, first=cut)
        assert np.array_equal(np.concatenate([a, b]), off)
        # sampled medians / keypoints against the oracle
        cfg = oracle.config(320, 224)
        rng = np.random.default_rng(0)
        reg.register(n)
        for i in rng.integers(0, n, size=12):
            omed, okps = oracle.extract(cfg, seq.frames[i])
            assert np.array_equal(reg.fetch_medians(1, first=int(i))[0], omed)
            kps = reg.keypoints(int(i))
            assert np.array_equal(kps["x"], okps["x"]) and np.array_equal(kps["code"], okps["code"])


def test_full_size_640x480_properties():
    """BASELINE config 4: 640x480, large scroll offsets, dense keypoints."""
    n = 600
    seq = synth.scrolling_tilemap(n, 640, 480, seed=4, speckle=0.10, vmax=(48, 48))
    with remap_b200.Registrar(640, 480, max_frames=n) as reg:
        reg.upload(seq.frames)
        off, _ = reg.register(n)
        ok = (off["flags"] & RB_OFFSET_VALID) != 0
        assert ok.all()
        assert np.array_equal(np.stack([off["dx"], off["dy"]], 1), seq.true_offsets)


def test_scene_cuts_start_new_fragments():
    """Config 5 flavour: hard cuts between levels must give 'no offset' exactly at the cuts."""
    n = 400
    seq = synth.scrolling_tilemap(n, 320, 224, seed=5, cut_every=40, levels=3)
    cuts = np.nonzero(seq.level[1:] != seq.level[:-1])[0]
    assert len(cuts) > 3
    with remap_b200.Registrar(320, 224, max_frames=n) as reg:
        reg.upload(seq.frames)
        off, _ = reg.register(n)
    valid = (off["flags"] & RB_OFFSET_VALID) != 0
    assert not valid[cuts].any()
    keep = np.ones(n - 1, bool); keep[cuts] = False
    assert valid[keep].all()
    assert np.array_equal(np.stack([off["dx"], off["dy"]], 1)[keep], seq.true_offsets[keep])


def test_foreground_mask_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "fgmask.npz"))
    bg = z["bg"]
    for k in range(int(z["n"])):
        px, py = (int(v) for v in z[f"pos{k}"])
        frame = z[f"frame{k}"]
        H, W = frame.shape
        if W / 4 <= 8 or H / 2 <= 8:
            continue  # smaller than any registrable window
        with remap_b200.Registrar(W, H, max_frames=2) as reg:
            assert np.array_equal(reg.foreground_mask(bg, px, py, frame), z[f"mask{k}"])


def test_foreground_mask_sprites_vs_oracle():
    """Config 3 flavour: sprites over a scrolling background; mask against the true background."""
    seq = synth.scrolling_tilemap(6, 320, 224, seed=6, sprites=10)
    clean = synth.scrolling_tilemap(6, 320, 224, seed=6, sprites=0)
    H, W = 224, 320
    x0, y0 = seq.path[:, 0].min(), seq.path[:, 1].min()
    x1, y1 = seq.path[:, 0].max() + W, seq.path[:, 1].max() + H
    bg = np.zeros((y1 - y0, x1 - x0), np.uint8)
    for i in range(6):
        px, py = seq.path[i] - (x0, y0)
        bg[py:py + H, px:px + W] = clean.frames[i]
    with remap_b200.Registrar(W, H, max_frames=8) as reg:
        reg.upload(seq.frames)
        for i in range(6):
            px, py = (int(v) for v in seq.path[i] - (x0, y0))
            want = oracle.foreground_mask(bg, px, py, seq.frames[i])
            assert np.array_equal(reg.foreground_mask(bg, px, py, seq.frames[i]), want)
            assert np.array_equal(reg.foreground_mask_resident(bg, px, py, i), want)
            assert (want == 0).any() and (want == 255).any()


def test_errors_are_reported_not_thrown():
    with remap_b200.Registrar(320, 224, max_frames=4) as reg:
        with pytest.raises(remap_b200.RemapError):
            reg.register(3)  # nothing uploaded
        with pytest.raises(remap_b200.RemapError):
            reg.upload(np.zeros((5, 224, 320), np.uint8))  # beyond capacity
    with pytest.raises(remap_b200.RemapError):
        remap_b200.Registrar(16, 16, max_frames=4)  # too small for the 4x2 grid


def _assemble_on_gpu(frames, pos):
    from remap_b200 import PLACEMENT_DTYPE, shard
    n, H, W = frames.shape
    zx, zy, mw, mh = shard.fragment_extents(pos, W, H)
    pl = np.zeros(n, PLACEMENT_DTYPE)
    pl["frame"] = np.arange(n)
    pl["x"] = pos[:, 0] - zx
    pl["y"] = pos[:, 1] - zy
    with remap_b200.Registrar(W, H, max_frames=max(n, 2)) as reg:
        reg.upload(frames)
        dots, image, mask = reg.blit_blend(pl, mw, mh)
    return (zx, zy), dots, image, mask


@pytest.mark.parametrize("case", ["scroll", "back_and_forth", "wrap"])
def test_map_assembly_blit_blend_matches_reference_semantics(case):
    """rb_blit_blend (gathered, no atomics) == fragment::blit frame by frame + blend, including the way the
    reference grows the map (frame-sized steps) and the uint16 wrap of the counters."""
    if case == "scroll":
        seq = synth.scrolling_tilemap(120, 320, 224, seed=21, vmax=(9, 7))
        frames, pos = seq.frames, shard_positions(seq)
    elif case == "back_and_forth":
        seq = synth.scrolling_tilemap(60, 200, 136, seed=22, vmax=(30, 20))
        frames = np.concatenate([seq.frames, seq.frames[::-1]])
        p = shard_positions(seq)
        pos = np.concatenate([p, p[::-1]])
        pos = pos - np.array([[150, -90]])  # negative and positive excursions around the first frame
        pos[0] = 0
    else:  # one small frame placed 65,540 times on the same spot: the counters wrap past 65,535
        rng = np.random.default_rng(23)
        frames = np.repeat(rng.integers(0, 16, size=(1, 40, 64), dtype=np.uint8), 70, axis=0)
        pos = np.zeros((70, 2), np.int64)
    zero, dots, image, mask = _assemble_on_gpu(frames, pos)
    if case == "wrap":
        # 70 placements are cheap to check exactly; the wrap itself needs > 65,535 visits: repeat the list
        from remap_b200 import PLACEMENT_DTYPE
        n_rep = 65540
        pl = np.zeros(n_rep, PLACEMENT_DTYPE)
        with remap_b200.Registrar(64, 40, max_frames=2) as reg:
            reg.upload(frames[:2])
            dots, image, mask = reg.blit_blend(pl, 64, 40)
        want = np.zeros((40, 64, 16), np.uint16)
        np.put_along_axis(want, frames[0][:, :, None].astype(np.int64), np.uint16(n_rep % 65536), axis=2)
        assert np.array_equal(dots, want)
        assert np.array_equal(image, frames[0]) and mask.all()  # 4 != 0: still the only non-zero bin
        return
    want = oracle.assemble_fragment(frames, pos)
    assert zero == want["zero"] and dots.shape == want["dots"].shape
    assert np.array_equal(dots, want["dots"])
    assert np.array_equal(image, want["image"]) and np.array_equal(mask, want["mask"])


def shard_positions(seq):
    """positions as frc::collector accumulates them from the true offsets (first frame at 0, 0)"""
    return np.concatenate([[[0, 0]], np.cumsum(seq.true_offsets, axis=0)]).astype(np.int64)


def test_cooperative_declare_kernel_equals_the_one_thread_per_pair_reference():
    """K3 as launched (one lane group per pair, shuffles) against rbm::declare_pair run one thread per pair
    (RB_K3_REFERENCE=1; the same function the CPU suite runs on the host against the oracle), on sequences
    where the vote is close: 16-px parallax bands (tie-sensitive pairs), scene cuts, sprites."""
    cases = [synth.scrolling_tilemap(400, 320, 224, seed=61, parallax=16),
             synth.scrolling_tilemap(300, 320, 224, seed=62, cut_every=25, levels=3, sprites=8),
             synth.scrolling_tilemap(200, 200, 136, seed=63, parallax=32, speckle=0.02)]
    for seq in cases:
        n, H, W = seq.frames.shape
        outs = []
        for ref in (False, True):
            if ref:
                os.environ["RB_K3_REFERENCE"] = "1"
            try:
                with remap_b200.Registrar(W, H, max_frames=n) as reg:
                    reg.upload(seq.frames)
                    off, _ = reg.register(n)
                    outs.append(off.copy())
            finally:
                os.environ.pop("RB_K3_REFERENCE", None)
        assert np.array_equal(outs[0], outs[1])
        flagged = (outs[0]["flags"] & RB_OFFSET_TIE_SENSITIVE) != 0
        valid = (outs[0]["flags"] & RB_OFFSET_VALID) != 0
        assert valid.any() and (~valid).any() or not flagged.any()


def test_aws_compare_matches_reference_and_oracle(golden_dir):
    """rb_aws_compare (aws::details::compare over a run of frames) against the real reference's heat maps and,
    on a screen-sized sequence, against the numpy restatement."""
    z = np.load(os.path.join(golden_dir, "awsheat.npz"))
    frames, ref = z["frames"], z["heat"]
    n, H, W = frames.shape
    with remap_b200.Registrar(W, H, max_frames=n) as reg:
        reg.upload(frames)
        heat, fc = reg.aws_compare(n)
        assert np.array_equal(heat, ref[-1])
        for k in range(len(ref)):
            assert np.array_equal((fc > k).astype(np.uint8), ref[k]), k
        h2, fc2 = reg.aws_compare(6, first=3, heat=ref[2])   # continue from a given map, sub-range
        assert np.array_equal(h2, ref[7])
    rng = np.random.default_rng(5)
    seq = synth.scrolling_tilemap(300, 320, 224, seed=8)
    screen = np.full((300, 312, 388), 14, np.uint8)          # the reference's screen size (src/main.cpp:194-244)
    screen[:, 40:264, 32:352] = seq.frames
    flick = rng.integers(0, 300, size=50)
    screen[flick, rng.integers(0, 312, size=50), rng.integers(0, 388, size=50)] = 3
    with remap_b200.Registrar(388, 312, max_frames=300) as reg:
        reg.upload(screen)
        heat, fc = reg.aws_compare(300)
    oh, ofc = oracle.aws_compare(screen)
    assert np.array_equal(heat, oh) and np.array_equal(fc, ofc)


def test_aws_compare_single_frame_and_identical_frames():
    frames = synth.random_frames(1, 96, 64, seed=9).repeat(5, axis=0)
    with remap_b200.Registrar(96, 64, max_frames=5) as reg:
        reg.upload(frames)
        heat, fc = reg.aws_compare(1)               # no pair at all
        assert heat.min() == 1 and (fc == 0xFFFFFFFF).all()
        heat, fc = reg.aws_compare(5)               # five identical frames: nothing ever changes
        assert heat.min() == 1 and (fc == 0xFFFFFFFF).all()


def test_back_to_back_host_registrations_without_a_fetch_in_between():
    """Two rb_register_host_async calls in a row (no synchronising call between them): the second must not pack into
    staging memory the first call's copies are still reading."""
    import torch
    n, W, H = 1200, 320, 224
    seq = synth.scrolling_tilemap(2 * n, W, H, seed=94)
    a = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    b = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    a.numpy()[...] = seq.frames[:n]
    b.numpy()[...] = seq.frames[n:]
    with remap_b200.Registrar(W, H, max_frames=2 * n, upload_chunk=64) as reg:
        reg.upload(seq.frames)
        want, _ = reg.register(2 * n)
    with remap_b200.Registrar(W, H, max_frames=2 * n, upload_chunk=64) as reg:
        reg.register_host_async(a.numpy(), first=0)
        reg.register_host_async(b.numpy(), first=n)
        got_b = reg.fetch_offsets(n - 1)
        reg.register_async(2 * n)  # the frames of both calls are resident and intact
        got = reg.fetch_offsets(2 * n - 1)
    assert np.array_equal(got_b, want[n:]) and np.array_equal(got, want)
