"""-m gpu: the collector-shaped C++ front of the C ABI (include/frc_b200.hpp) is a drop-in for the
reference's frc::collector.  oracle/_ref/shim_harness (built by oracle/build_ref.py from the
reference's own headers + our header, linked against libremap_b200.so) runs both collectors on
the same frames with the reference's feeder concept, nic::compress and a recording callback and
compares fragments, dot histograms, frame positions, compressed images AND compressed medians, the
callback sequence, and (fill_keys) the kpr::grid handed to the callback."""
import os
import subprocess

import numpy as np
import pytest

from remap_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "shim_harness")

pytestmark = pytest.mark.gpu


def _run(frames, batch, fill, tmp_path, gpu_blit=False, filter=False, devices=None):
    if not os.path.exists(SHIM):
        pytest.fail("oracle/_ref/shim_harness missing: run `python oracle/build_ref.py` in the build container")
    n, H, W = frames.shape
    path = os.path.join(tmp_path, "frames.bin")
    np.ascontiguousarray(frames, np.uint8).tofile(path)
    cmd = [SHIM, path, str(W), str(H), str(n), str(batch), str(int(fill)), str(int(gpu_blit)), str(int(filter))]
    if devices:
        cmd += ["0", ",".join(str(d) for d in devices)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    assert "\nIDENTICAL" in "\n" + r.stdout, r.stdout
    return r.stdout[r.stdout.index("IDENTICAL"):] if not filter else r.stdout


@pytest.mark.parametrize("batch", [2, 7, 64])
def test_collector_shim_matches_reference_collector(batch, tmp_path):
    seq = synth.scrolling_tilemap(40, 320, 224, seed=11)
    out = _run(seq.frames, batch, True, str(tmp_path))
    assert "1 fragments, 40 frames, 39 callbacks" in out, out


def test_collector_shim_scene_cuts_open_fragments(tmp_path):
    seq = synth.scrolling_tilemap(60, 320, 224, seed=5, cut_every=17)
    out = _run(seq.frames, 16, False, str(tmp_path))
    nfrag = int(out.split()[1])
    assert nfrag >= 2 and "60 frames" in out, out


def test_collector_shim_odd_size_and_long_run(tmp_path):
    seq = synth.scrolling_tilemap(300, 323, 227, seed=9)
    _run(seq.frames, 128, False, str(tmp_path))


@pytest.mark.parametrize("batch", [5, 64])
def test_collector_shim_with_gpu_map_assembly(batch, tmp_path):
    """options::gpu_blit: the fragments' dot maps come from rb_blit_blend instead of fragment::blit on the host
    and must be identical to the reference collector's, fragment for fragment (zero, size, every histogram)."""
    seq = synth.scrolling_tilemap(120, 320, 224, seed=13, cut_every=45, vmax=(9, 7))
    out = _run(seq.frames, batch, False, str(tmp_path), gpu_blit=True)
    assert "dots from rb_blit_blend" in out and int(out.split()[1]) >= 2, out


def test_collector_shim_gpu_blit_with_keys_across_batches(tmp_path):
    """gpu_blit + fill_keys with a batch smaller than the sequence: from the second batch on a frame's store slot
    differs from its index in the batch; the kpr::grid handed to the callback must still be that frame's."""
    seq = synth.scrolling_tilemap(40, 320, 224, seed=17)
    out = _run(seq.frames, 9, True, str(tmp_path), gpu_blit=True)
    assert "keys compared" in out and "dots from rb_blit_blend" in out, out


@pytest.mark.parametrize("members,batch", [(2, 64), (3, 25), (8, 200)])
def test_collector_shim_over_several_devices(members, batch, tmp_path):
    """options::devices: every batch is split over the members of an rb_group (one context per device; on a one-GPU
    box they share cuda:0) -- fragments, positions, compressed medians, callbacks and keys as the reference's."""
    import torch
    ngpu = torch.cuda.device_count()
    seq = synth.scrolling_tilemap(130, 320, 224, seed=19, cut_every=50)
    out = _run(seq.frames, batch, True, str(tmp_path), devices=[i % ngpu for i in range(members)])
    assert "130 frames" in out and "keys compared" in out, out


def test_filter_shim_matches_reference_fdf_filter(tmp_path):
    """include/fdf_b200.hpp (rb_filter_fragment) against the reference's fdf::filter on the reference
    collector's own fragments: filtered dots, frame lists and every fde::mask handed to the callback."""
    seq = synth.scrolling_tilemap(90, 320, 224, seed=21, sprites=6, cut_every=40)
    out = _run(seq.frames, 32, False, str(tmp_path), filter=True)
    assert "FILTER IDENTICAL" in out and "90 masks" in out, out


def test_filter_shim_resident_mode(tmp_path):
    """Pass 2 on the frames frc_b200::collector (gpu_blit) left on the device -- no decompression, no upload --
    gives the reference's filtered dots too."""
    seq = synth.scrolling_tilemap(90, 320, 224, seed=22, sprites=5, cut_every=40)
    out = _run(seq.frames, 32, False, str(tmp_path), gpu_blit=True, filter=True)
    assert "FILTER IDENTICAL" in out and "RESIDENT FILTER IDENTICAL" in out, out


def test_lean_collector_and_resident_filter(tmp_path):
    """gpu_blit = 2: the collector keeps no compressed copies and fetches no medians; its fragments (dots from
    rb_blit_blend) still equal the reference's, and pass 2 runs on ITS fragments in resident mode."""
    seq = synth.scrolling_tilemap(90, 320, 224, seed=23, sprites=5, cut_every=40)
    out = _run(seq.frames, 32, False, str(tmp_path), gpu_blit=2, filter=True)
    assert "RESIDENT FILTER IDENTICAL" in out and "\nIDENTICAL" in "\n" + out, out
