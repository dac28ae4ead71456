"""Shared parity checks: compare an implementation's intermediates with the REAL reference's dump
(oracle/refdump.py -> oracle/_ref/ref_harness) or with committed golden dumps of it.

An "implementation" here is anything that yields, per frame, (median, keypoints) and, per pair,
(result, per-region votes, per-region histograms) in the oracle's record layouts
(oracle/oracle.py dtypes) -- the C restatement and the CUDA path both do.
"""
from __future__ import annotations

import numpy as np

NREG = 8


def kps_by_region(kps, nreg=NREG):
    """structured keypoints (x, y, code, region_mask) -> per region sorted (x, y, code) arrays."""
    out = []
    for r in range(nreg):
        sel = kps[(kps["region_mask"] >> r) & 1 == 1]
        order = np.lexsort((sel["y"], sel["x"]))
        out.append(sel[order])
    return out


def check_frame(ref_frame, median, kps, where=""):
    """ref_frame: dict(median, regions[...]) from the reference dump."""
    assert np.array_equal(ref_frame["median"], median), f"{where}: median differs"
    per_region = kps_by_region(kps)
    for r in range(NREG):
        rr = ref_frame["regions"][r]
        mine = per_region[r]
        assert rr["n"] == len(mine), f"{where} region {r}: {rr['n']} keypoints in reference, {len(mine)} here"
        assert np.array_equal(rr["kps"]["x"], mine["x"]) and np.array_equal(rr["kps"]["y"], mine["y"]), \
            f"{where} region {r}: keypoint coordinates differ"
        assert np.array_equal(rr["kps"]["code"], mine["code"]), f"{where} region {r}: 13-byte codes differ"
        w = mine["code"][:, 12] & 0xF
        assert rr["w1"] == int((w == 1).sum()) and rr["w2"] == int((w == 2).sum()), f"{where} region {r}: weights"


def check_pair(ref_pair, result, votes, bins_per_region, where=""):
    """-> 'flagged' | 'ok'.  Raises on any mismatch that is not attributable to the reference's
    undefined tie order."""
    assert ref_pair["active"] == int(result["active"]), f"{where}: active"
    for r in range(NREG):
        rr = ref_pair["regions"][r]
        v = votes[r]
        assert rr["use_all"] == bool(v["use_all"]), f"{where} region {r}: weight switch"
        if bins_per_region is not None:
            b = bins_per_region[r]
            assert len(rr["bins"]) == len(b), f"{where} region {r}: {len(rr['bins'])} bins vs {len(b)}"
            assert np.array_equal(rr["bins"]["dx"], b["dx"]) and np.array_equal(rr["bins"]["dy"], b["dy"]) \
                and np.array_equal(rr["bins"]["cnt"], b["cnt"]), f"{where} region {r}: histogram differs"
        assert len(rr["bins"]) == int(v["nbins"]), f"{where} region {r}: nbins"
        # tickets: the COUNT sequence is tie-independent and must match exactly; the offsets must
        # match wherever that count is unique in the histogram
        nt = int(v["nticket"])
        assert len(rr["ticket"]) == nt, f"{where} region {r}: ticket size"
        cnts = rr["bins"]["cnt"]
        for k in range(nt):
            assert int(rr["ticket"]["cnt"][k]) == int(v["ticket"]["cnt"][k]), f"{where} region {r}: ticket count {k}"
            c = int(rr["ticket"]["cnt"][k])
            if int((cnts == c).sum()) == 1:
                assert (int(rr["ticket"]["dx"][k]), int(rr["ticket"]["dy"][k])) == \
                    (int(v["ticket"]["dx"][k]), int(v["ticket"]["dy"][k])), f"{where} region {r}: ticket offset {k}"
            assert int(v["ngt"][k]) == int((cnts > c).sum()) and int(v["nge"][k]) == int((cnts >= c).sum()), \
                f"{where} region {r}: tie statistics {k}"
    if result["tie_sensitive"]:
        return "flagged"
    assert ref_pair["valid"] == bool(result["valid"]), f"{where}: declared validity differs on an unflagged pair"
    if ref_pair["valid"]:
        assert (ref_pair["dx"], ref_pair["dy"]) == (int(result["dx"]), int(result["dy"])), \
            f"{where}: declared offset differs on an unflagged pair"
    return "ok"


def positions_from_results(valid, dx, dy):
    """frc loop (src/frc.hpp:109-115): position += off, or new fragment at (0, 0)."""
    n = len(valid) + 1
    pos = np.zeros((n, 3), np.int32)
    f = x = y = 0
    for i in range(1, n):
        if valid[i - 1]:
            x += int(dx[i - 1]); y += int(dy[i - 1])
        else:
            f += 1; x = 0; y = 0
        pos[i] = (f, x, y)
    return pos
