"""Helpers shared by the -m gpu parity tests: adapt Registrar outputs to tests/parity.py records."""
import numpy as np

import remap_b200
from remap_b200 import RB_OFFSET_TIE_SENSITIVE, RB_OFFSET_VALID


def result_record(offset, ballots):
    return dict(valid=bool(offset["flags"] & RB_OFFSET_VALID), dx=int(offset["dx"]), dy=int(offset["dy"]),
                tie_sensitive=bool(offset["flags"] & RB_OFFSET_TIE_SENSITIVE),
                active=int((ballots["n_curr"] > 0).sum()))


def run_sequence(frames, want_medians=True, **kw):
    """Registers a whole sequence on cuda:0 through the C ABI.
    -> dict(reg, offsets, medians) ; caller closes reg."""
    n, H, W = frames.shape
    reg = remap_b200.Registrar(W, H, max_frames=max(n, 2), compute_median=True, **kw)
    reg.upload(frames)
    offsets, medians = reg.register(n, want_medians=want_medians)
    return dict(reg=reg, offsets=offsets, medians=medians)


def compare_with_oracle(frames, oracle, name, taps=((0, 0), (0, 5)), **kw):
    """Full diff of the CUDA path against the C restatement on the same frames."""
    n, H, W = frames.shape
    cfg = oracle.config(W, H)
    out = run_sequence(frames, **kw)
    reg = out["reg"]
    try:
        prev = None
        flagged = 0
        for i in range(n):
            omed, okps = oracle.extract(cfg, frames[i])
            assert np.array_equal(omed, out["medians"][i]), f"{name} frame {i}: median"
            kps = reg.keypoints(i)
            assert len(kps) == len(okps), f"{name} frame {i}: {len(kps)} keypoints vs oracle {len(okps)}"
            for fld in ("x", "y", "code", "weight", "region_mask"):
                assert np.array_equal(kps[fld], okps[fld]), f"{name} frame {i}: keypoint field {fld}"
            if i > 0:
                ores, ovotes = oracle.match(cfg, prev, okps)
                ballots = reg.region_ballots(i - 1)
                for fld in ballots.dtype.names:
                    assert np.array_equal(ballots[fld], ovotes[fld]), f"{name} pair {i}: ballot field {fld}"
                rec = result_record(out["offsets"][i - 1], ballots)
                assert rec["valid"] == bool(ores["valid"]) and rec["tie_sensitive"] == bool(ores["tie_sensitive"]), \
                    f"{name} pair {i}: flags {rec} vs {ores}"
                assert (rec["dx"], rec["dy"]) == (int(ores["dx"]), int(ores["dy"])), f"{name} pair {i}: offset"
                assert rec["active"] == int(ores["active"])
                flagged += rec["tie_sensitive"]
                for (tp, tr) in taps:
                    if tp == i - 1:
                        bins = reg.region_votes(tp, tr)
                        obins = oracle.region_bins(cfg, prev, okps, tr)
                        assert np.array_equal(bins, obins), f"{name} pair {i} region {tr}: histogram"
            prev = okps
        return flagged
    finally:
        reg.close()
