"""CPU: the full-sequence digest of the real reference (oracle/_ref/ref_harness digest) is pinned against the C
restatement -- every digest definition (median, grid insertions, offset histograms) is recomputed in numpy from
the restatement's outputs -- and its sharded runs (any thread count) agree with a single-threaded one, including the
stitched frc::collector positions."""
import numpy as np
import pytest

import digest_check
from oracle import oracle, refdump
from remap_b200 import synth

pytestmark = pytest.mark.skipif(not digest_check.have_ref(), reason="oracle/_ref/ref_harness not built")


def test_reference_digest_equals_restatement():
    n, W, H = 14, 320, 224
    seq = synth.scrolling_tilemap(n, W, H, seed=61, cut_every=5)
    ref = digest_check.ref_digest(seq.frames, threads=3)
    cfg = oracle.config(W, H)
    prev = None
    for i in range(n):
        med, kps = oracle.extract(cfg, seq.frames[i])
        assert digest_check.median_hash(med) == int(ref["frames"]["median_hash"][i]), i
        kh, ins = digest_check.kp_hash(kps)
        assert kh == int(ref["frames"]["kp_hash"][i]) and ins == int(ref["frames"]["insertions"][i]), i
        if i > 0:
            res, votes = oracle.match(cfg, prev, kps)
            rp = ref["pairs"][i - 1]
            assert np.array_equal(rp["r"]["hist_hash"], votes["hist_hash"]), i
            assert np.array_equal(rp["r"]["nbins"], votes["nbins"]) and np.array_equal(rp["r"]["use_all"], votes["use_all"])
            assert np.array_equal(rp["r"]["tcnt"], votes["ticket"]["cnt"][:, :3])
            assert np.array_equal(ref["frames"]["n"][i], votes["n_curr"]) and np.array_equal(ref["frames"]["w2"][i], votes["w2_curr"])
            if not res["tie_sensitive"]:
                assert bool(rp["valid"]) == bool(res["valid"])
                if res["valid"]:
                    assert (int(rp["dx"]), int(rp["dy"])) == (int(res["dx"]), int(res["dy"]))
        prev = kps
    assert (~(ref["pairs"]["valid"] != 0)).sum() >= 1, "the sequence should hold a scene cut"


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_sharded_digest_and_stitched_collector_positions(threads):
    n, W, H = 23, 200, 136
    seq = synth.scrolling_tilemap(n, W, H, seed=62, cut_every=6, world_w=512, world_h=400)
    ref = digest_check.ref_digest(seq.frames, threads=threads)
    one = digest_check.ref_digest(seq.frames, threads=1)
    for k in ("frames", "pairs", "positions"):
        assert np.array_equal(ref[k], one[k]), k
    dump = refdump.ref_dump(seq.frames)
    assert np.array_equal(ref["positions"], dump["positions"])
    assert len(np.unique(ref["positions"][:, 0])) >= 2


def test_pairs_mode_equals_digest_declarations(tmp_path):
    """ref_harness pairs (what tools/resolve_flagged.py replays flagged pairs with) declares what the digest's kpm::match
    declared for the same pairs."""
    import os
    import subprocess
    n, W, H = 12, 320, 224
    seq = synth.scrolling_tilemap(n, W, H, seed=63, cut_every=4)
    ref = digest_check.ref_digest(seq.frames, threads=2, collector=False)
    fin, pin, pout = (os.path.join(tmp_path, x) for x in ("f.bin", "p.bin", "o.bin"))
    seq.frames.tofile(fin)
    pairs = np.array([0, 3, 4, 7, 10], np.uint32)
    pairs.tofile(pin)
    subprocess.check_call([digest_check.REF_BIN, "pairs", fin, str(W), str(H), str(n), pin, pout])
    res = np.fromfile(pout, "<i4").reshape(-1, 3)
    rp = ref["pairs"][pairs]
    assert np.array_equal(res[:, 0] != 0, rp["valid"] != 0)
    assert np.array_equal(res[:, 1], rp["dx"]) and np.array_equal(res[:, 2], rp["dy"])
