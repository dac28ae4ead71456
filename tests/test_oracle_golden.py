"""CPU: the C restatement (oracle/remap_oracle.c) against committed dumps of the REAL reference.

This is the oracle's pin (the reference has no test vectors of its own)."""
import glob
import os

import numpy as np
import pytest

import parity
from oracle import oracle, refdump

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
               if not p.endswith(("fgmask.npz", "awsheat.npz")) and not os.path.basename(p).startswith(("filter_", "splice_")))


def test_luts_known_answer():
    # values the reference's consteval LUT generator produces (src/cpl.hpp:163-217), SURVEY.md a1
    n2o, o2n = oracle.luts()
    assert n2o.tolist() == [0, 15, 2, 12, 6, 9, 3, 13, 5, 1, 7, 4, 8, 14, 10, 11]
    assert o2n.tolist() == [0, 9, 2, 6, 11, 8, 4, 10, 12, 5, 14, 15, 3, 7, 13, 1]


def _runs(a):
    out, s = [], 0
    for i in range(1, len(a) + 1):
        if i == len(a) or a[i] != a[s]:
            out.append((s, i, int(a[s])))
            s = i
    return out


def test_sections_known_answer():
    # SURVEY.md Appendix A.4 (derived from src/kpe.hpp:157-192,235-277 for 320x224 and 640x480)
    cs, rs = oracle.sections(oracle.config(320, 224))
    assert _runs(cs) == [(0, 2, 0), (2, 74, 1), (74, 90, 3), (90, 162, 2), (162, 178, 6), (178, 250, 4),
                         (250, 266, 12), (266, 318, 8), (318, 320, 0)]
    assert _runs(rs) == [(0, 2, 0), (2, 107, 1), (107, 123, 3), (123, 220, 2), (220, 224, 0)]
    cs, rs = oracle.sections(oracle.config(640, 480))
    assert [r[:2] for r in _runs(cs)] == [(0, 2), (2, 154), (154, 170), (170, 322), (322, 338), (338, 490),
                                          (490, 506), (506, 638), (638, 640)]
    assert [r[:2] for r in _runs(rs)] == [(0, 2), (2, 235), (235, 251), (251, 476), (476, 480)]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_dump(name, golden_dir):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    frames = z["frames"]
    ref = refdump.parse_dump(z["dump"].tobytes())
    N, H, W = frames.shape
    cfg = oracle.config(W, H)
    prev = None
    valid, dx, dy = [], [], []
    flagged = 0
    for i in range(N):
        med, kps = oracle.extract(cfg, frames[i])
        parity.check_frame(ref["frames"][i], med, kps, f"{name} frame {i}")
        if i > 0:
            res, votes = oracle.match(cfg, prev, kps)
            bins = [oracle.region_bins(cfg, prev, kps, r) for r in range(8)]
            st = parity.check_pair(ref["pairs"][i - 1], res, votes, bins, f"{name} pair {i}")
            flagged += st == "flagged"
            valid.append(bool(res["valid"])); dx.append(int(res["dx"])); dy.append(int(res["dy"]))
        prev = kps
    if flagged == 0:
        # the unmodified frc::collector loop's (fragment, x, y) per frame
        assert np.array_equal(parity.positions_from_results(valid, dx, dy), ref["positions"])
    if name != "parallax":
        assert flagged == 0
    # whole-sequence entry point agrees with the per-frame calls
    out = oracle.register(cfg, frames, want_medians=True)
    assert [bool(v) for v in out["results"]["valid"]] == valid
    assert out["results"]["dx"].tolist() == dx and out["results"]["dy"].tolist() == dy
    for i in range(N):
        assert np.array_equal(out["medians"][i], ref["frames"][i]["median"])


def test_foreground_mask_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "fgmask.npz"))
    bg = z["bg"]
    for k in range(int(z["n"])):
        px, py = (int(v) for v in z[f"pos{k}"])
        m = oracle.foreground_mask(bg, px, py, z[f"frame{k}"])
        assert np.array_equal(m, z[f"mask{k}"])
        assert set(np.unique(m).tolist()) <= {0, 255}


# ---- pass 2: fdf::filter (src/fdf.hpp:40-91) ---------------------------------------------------------
FILTER_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "filter_*.npz")))


def load_filter_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    return z["frames"], refdump.parse_filter_dump(z["dump"].tobytes())


@pytest.mark.parametrize("name", FILTER_CASES)
def test_oracle_foreground_matches_reference_filter(name, golden_dir):
    """The C restatement of fde::extractor::extract + fde::mask (and the masked blit in numpy) against the
    real reference's frc::collector + fdf::filter run: every frame's contours and mask, every fragment's dots."""
    frames, ref = load_filter_case(golden_dir, name)
    N, H, W = frames.shape
    cfg = oracle.config(W, H)
    medians = np.stack([oracle.extract(cfg, f)[0] for f in frames])
    assert len(ref["frames"]) == N and FILTER_CASES
    by_frag = {}
    for rec in ref["frames"]:
        by_frag.setdefault(rec["fragment"], []).append(rec)
    for fi, recs in by_frag.items():
        bg = ref["backgrounds"][fi]["image"]
        for rec in recs:
            i = rec["number"]
            mask, cont = oracle.foreground(bg, rec["x"], rec["y"], frames[i], medians[i])
            assert np.array_equal(mask, rec["mask"]), f"{name} frame {i}: mask"
            assert len(cont) == len(rec["contours"]), f"{name} frame {i}: contour count"
            for fld in cont.dtype.names:
                assert np.array_equal(cont[fld], rec["contours"][fld]), f"{name} frame {i}: contour field {fld}"
        idx = [r["number"] for r in recs]
        pos = np.array([[r["x"], r["y"]] for r in recs])
        mh, mw = bg.shape
        out = oracle.filter_fragment(frames[idx], medians[idx], pos, mw, mh)
        assert np.array_equal(out["background"], bg), f"{name} fragment {fi}: background"
        assert np.array_equal(out["dots"], ref["fragments"][fi]["dots"]), f"{name} fragment {fi}: dots"


# ---- fragment splicing: fgs::splice (src/fgs.hpp:187-213) ------------------------------------------------
SPLICE_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "splice_*.npz")))


def check_cell_match(got, want, what):
    """got: record with the fields of oracle.CELL_MATCH_DTYPE; want: a match of refdump.parse_splice_dump."""
    assert int(got["offsets"]) == want["offsets"] and int(got["pairs"]) == want["pairs"], what
    assert int(got["ties"]) == want["ties"], what
    if want["offsets"] == 0:
        assert not got["valid"] and not want["valid"], what
        return
    assert int(got["matched_keypoints"]) == want["matched_keypoints"], what
    if want["ties"] > 1:
        return  # the reference's pick among tied offsets is std::unordered_map iteration order: not comparable
    assert (int(got["dx"]), int(got["dy"])) == want["best"], what
    assert int(got["matched_cells"]) == want["matched_cells"] and int(got["active_cells"]) == want["active_cells"], what
    assert bool(got["valid"]) == want["valid"], what
    if want["valid"]:
        assert (int(got["dx"]), int(got["dy"])) == want["vote"] and int(got["matched_keypoints"]) == want["count"], what


@pytest.mark.parametrize("name", SPLICE_CASES)
def test_oracle_cell_match_matches_reference(name, golden_dir):
    """The C restatement of fgs's snippets (blend + 1x1-grid kpe) and of the cellular kpm::match against the
    real reference's dump of every snippet and every snippet pair."""
    z = np.load(os.path.join(golden_dir, f"{name}.npz"))
    ref = refdump.parse_splice_dump(z["dump"].tobytes())
    assert SPLICE_CASES and len(ref["fragments"]) >= 2
    snips = [oracle.snippet(f["dots"]) for f in ref["fragments"]]
    for s, r in zip(snips, ref["snippets"]):
        order = np.lexsort((s["kps"]["x"], s["kps"]["y"]))
        assert np.array_equal(s["mask"], r["mask"])
        for fld in ("x", "y", "code"):
            assert np.array_equal(s["kps"][fld][order], r["kps"][fld]), fld
    for m in ref["matches"]:
        check_cell_match(oracle.cell_match(snips[m["prev"]], snips[m["curr"]]), m, f"{name} {m['prev']}-{m['curr']}")


def test_oracle_aws_compare_matches_reference(golden_dir):
    """numpy restatement of aws::details::compare against the real reference's heat map after every pair."""
    z = np.load(os.path.join(golden_dir, "awsheat.npz"))
    frames, ref = z["frames"], z["heat"]
    heat, fc = oracle.aws_compare(frames)
    assert np.array_equal(heat, ref[-1])
    for k in range(len(ref)):  # the map after pair k is "no pair up to k differed"
        assert np.array_equal((fc > k).astype(np.uint8), ref[k]), k
    assert ref[-1][2, 3] == 0 and fc[2, 3] == 4 and ref[-1][0, 0] == 1
