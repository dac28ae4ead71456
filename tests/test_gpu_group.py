"""-m gpu: rb_group -- several devices of one process behind the C ABI (the seam mpb::builder::collect needs,
src/mpb.hpp:52-61).  Contiguous frame ranges with a one-frame overlap, registered concurrently, pair results gathered
on the lead device: must equal one context over the whole sequence.  On a one-GPU box the members share cuda:0."""
import numpy as np
import pytest
import torch

import remap_b200
from remap_b200 import synth

pytestmark = pytest.mark.gpu


def _devices(k):
    n = torch.cuda.device_count()
    return [i % n for i in range(k)]


@pytest.mark.parametrize("members,n", [(1, 40), (2, 41), (3, 100), (8, 67), (5, 3), (4, 2)])
def test_group_equals_single_context(members, n):
    W, H = 320, 224
    seq = synth.scrolling_tilemap(n, W, H, seed=31, cut_every=23)
    with remap_b200.Registrar(W, H, max_frames=n) as reg:
        reg.upload(seq.frames)
        off_a, med_a = reg.register(n, want_medians=True)
    pinned = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[...] = seq.frames
    with remap_b200.Group(W, H, n, _devices(members), upload_chunk=16) as grp:
        assert len(grp) == members
        for _ in range(2):
            off_b = grp.register_host(pinned.numpy())
        med_b = grp.fetch_medians(n)
        covered = np.zeros(n, np.int32)
        for i in range(members):
            first, end, own = grp.range(i)
            covered[own:end] += 1
            if end - first >= 1:  # the member's own taps work on its slots
                m = grp.member(i)
                kp = m.keypoints(end - first - 1)
                assert len(kp) > 0
        assert (covered == 1).all(), "every frame is owned by exactly one member"
    assert np.array_equal(off_a, off_b)
    assert np.array_equal(med_a, med_b)


def test_group_rejects_what_it_cannot_hold():
    with remap_b200.Group(320, 224, 10, _devices(2)) as grp:
        frames = np.zeros((11, 224, 320), np.uint8)
        with pytest.raises(remap_b200.RemapError):
            grp.register_host(frames)
