"""-m gpu: FULL-SEQUENCE parity against the real reference (not a sample, not the restatement): every frame's
median image and grid, every pair's per-region histogram, ticket and declared offset, and the frc::collector
positions, through the digests of tests/digest_check.py.  Sizes: BASELINE configs[1] whole (20,000 frames), the
others at one-GPU sizes the host's reference finishes in seconds (tools/run_config.py does them at full size)."""
import json
import os

import numpy as np
import pytest

import digest_check
import remap_b200
from remap_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not digest_check.have_ref(), reason="oracle/_ref/ref_harness not built")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(name, **kw):
    path = os.path.join(ROOT, "gpurun_out", "digest_parity.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = kw
    json.dump(data, open(path, "w"), indent=1)


def _run(name, seq, host_path=False, **kw):
    n, H, W = seq.frames.shape
    ref = digest_check.ref_digest(seq.frames)
    with remap_b200.Registrar(W, H, max_frames=n, **kw) as reg:
        if host_path:
            reg.register_host_async(seq.frames)
            off = reg.fetch_offsets(n - 1)
        else:
            reg.upload(seq.frames)
            off, _ = reg.register(n)
        res = digest_check.compare(ref, digest_check.gpu_digest(reg, n, offsets=off))
        res["deferred_ballots"] = reg.deferred_count
    res["reference_seconds"] = ref["seconds"]
    _record(name, **{k: v for k, v in res.items()})
    assert res["mismatches"] == 0, res["mismatch_detail"]
    return res


def test_config2_every_frame_and_pair_against_the_reference():
    res = _run("config2_20000", synth.scrolling_tilemap(20000, 320, 224, seed=1))
    assert res["pairs_compared"] == 19999 and res["reference_nullopt"] == 0


def test_config2_through_the_host_path():
    _run("config2_host_path_6000", synth.scrolling_tilemap(6000, 320, 224, seed=1), host_path=True)


def test_config3_sprites():
    _run("config3_sprites_6000", synth.scrolling_tilemap(6000, 320, 224, seed=3, sprites=12))


def test_config4_640x480():
    res = _run("config4_640x480_1000", synth.scrolling_tilemap(1000, 640, 480, seed=4, speckle=0.10, vmax=(48, 48)))
    assert res["deferred_ballots"] == 0


def test_config5_cuts_and_parallax():
    res = _run("config5_cuts_parallax_8000", synth.scrolling_tilemap(8000, 320, 224, seed=5, cut_every=2000, levels=3, parallax=32))
    assert res["reference_nullopt"] > 0


def test_heavy_ties_16px_parallax():
    """16-px parallax bands: the workload where SURVEY.md App. C saw declared offsets depend on the tie order."""
    _run("parallax16_3000", synth.scrolling_tilemap(3000, 320, 224, seed=6, parallax=16))


def test_flagged_pairs_resolved_offline_equal_the_reference_everywhere():
    """tools/resolve_flagged.py replays the flagged pairs through the compiled reference: afterwards every declared
    offset of the sequence equals the reference's (the flag is conservative and complete)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import resolve_flagged
    from remap_b200 import RB_OFFSET_VALID
    n, W, H = 1500, 320, 224
    seq = synth.scrolling_tilemap(n, W, H, seed=6, parallax=16)
    ref = digest_check.ref_digest(seq.frames, collector=False)
    with remap_b200.Registrar(W, H, max_frames=n) as reg:
        reg.upload(seq.frames)
        off, _ = reg.register(n)
    patched, info = resolve_flagged.resolve(seq.frames, off)
    assert info["flagged"] > 0, "the workload should hold tie-sensitive pairs"
    rvalid = ref["pairs"]["valid"] != 0
    assert np.array_equal((patched["flags"] & RB_OFFSET_VALID) != 0, rvalid)
    assert np.array_equal(patched["dx"][rvalid], ref["pairs"]["dx"][rvalid])
    assert np.array_equal(patched["dy"][rvalid], ref["pairs"]["dy"][rvalid])
    _record("parallax16_resolved_1500", **info)
