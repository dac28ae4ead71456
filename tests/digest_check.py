"""Full-sequence parity against the REAL reference: every frame and every pair, not a sample.

TEST INFRASTRUCTURE (used by tests/ and the tools/ runners; never by the product path).

oracle/_ref/ref_harness digest runs the reference's own kpe::extractor::extract, kpm::match, count_offsets and
top_offsets over a whole sequence on all host threads and writes, per frame, 64-bit digests of the median image
and of the grid's insertions, and per pair the declared offset plus, per region, the weight switch, the number of
histogram bins, a digest of the whole histogram and the ticket in the reference's own (libstdc++) order; optionally
the (fragment, x, y) of every frame from the unmodified frc::collector.  The CUDA path yields the same quantities
through rb_frame_digests / rb_fetch_ballots / rb_fetch_offsets; compare() diffs them one by one.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

REF_FRAME_DTYPE = np.dtype([("median_hash", "<u8"), ("kp_hash", "<u8"), ("insertions", "<u4"), ("n", "<u4", (8,)),
                            ("w2", "<u4", (8,))])
REF_REGION_DTYPE = np.dtype([("use_all", "<u4"), ("nbins", "<u4"), ("hist_hash", "<u4"), ("nticket", "<u4"),
                             ("tdx", "<i4", (3,)), ("tdy", "<i4", (3,)), ("tcnt", "<u4", (3,))])
REF_PAIR_DTYPE = np.dtype([("valid", "<u4"), ("dx", "<i4"), ("dy", "<i4"), ("active", "<u4"), ("r", REF_REGION_DTYPE, (8,))])
assert REF_FRAME_DTYPE.itemsize == 84 and REF_PAIR_DTYPE.itemsize == 16 + 8 * 52


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def ref_digest(frames: np.ndarray, threads: int | None = None, collector: bool = True, tmpdir: str | None = None) -> dict:
    """Runs the compiled reference in digest mode on (N, H, W) uint8 frames."""
    n, H, W = frames.shape
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory(dir=tmpdir) as td:
        fin, fout = os.path.join(td, "frames.bin"), os.path.join(td, "digest.bin")
        np.ascontiguousarray(frames, np.uint8).tofile(fin)
        r = subprocess.run([REF_BIN, "digest", fin, str(W), str(H), str(n), fout, str(threads), "1" if collector else "0"],
                           capture_output=True, text=True, timeout=7200)
        assert r.returncode == 0, (r.returncode, r.stderr[-400:])
        info = json.loads(r.stdout.strip().splitlines()[-1])
        buf = open(fout, "rb").read()
    assert buf[:4] == b"RMDG"
    w, h, nn, has_pos = np.frombuffer(buf, "<u4", 4, 4)
    assert (w, h, nn) == (W, H, n)
    pos = 20
    fr = np.frombuffer(buf, REF_FRAME_DTYPE, n, pos).copy()
    pos += n * REF_FRAME_DTYPE.itemsize
    pr = np.frombuffer(buf, REF_PAIR_DTYPE, n - 1, pos).copy()
    pos += (n - 1) * REF_PAIR_DTYPE.itemsize
    positions = None
    if has_pos:
        positions = np.frombuffer(buf, "<i4", 3 * n, pos).reshape(n, 3).copy()
        pos += 12 * n
    assert pos == len(buf)
    return dict(frames=fr, pairs=pr, positions=positions, seconds=info["seconds"], threads=info["threads"])


def gpu_digest(reg, n: int, first: int = 0, offsets=None) -> dict:
    """The same quantities from a Registrar whose last registration covered frames [first, first + n)."""
    return dict(frames=reg.frame_digests(n, first=first), ballots=reg.fetch_ballots(n - 1),
                offsets=offsets if offsets is not None else reg.fetch_offsets(n - 1))


def compare(ref: dict, gpu: dict, check_positions: bool = True) -> dict:
    """Diffs every frame and pair.  -> counts; `mismatches` must be 0 -- declared offsets may differ from the
    reference's only on pairs the CUDA path FLAGGED as tie-sensitive (std::unordered_map order decides them in the
    reference, src/kpm.hpp:127-159); those are counted in flagged_and_different."""
    from remap_b200 import RB_OFFSET_TIE_SENSITIVE, RB_OFFSET_VALID
    rf, rp, gf, gb, go = ref["frames"], ref["pairs"], gpu["frames"], gpu["ballots"], gpu["offsets"]
    n = len(rf)
    assert len(gf) == n and len(gb) == n - 1 and len(go) == n - 1
    bad = {}

    def note(name, mask):
        c = int(np.count_nonzero(mask))
        if c:
            bad[name] = dict(count=c, first=int(np.argmax(mask.reshape(len(mask), -1).any(axis=1))))

    note("median_hash", rf["median_hash"] != gf["median_hash"])
    note("kp_hash", rf["kp_hash"] != gf["kp_hash"])
    note("insertions", rf["insertions"] != gf["insertions"])
    if n > 1:
        note("n_curr", rf["n"][1:] != gb["n_curr"])
        note("w2_curr", rf["w2"][1:] != gb["w2_curr"])
        note("n_prev", rf["n"][:-1] != gb["n_prev"])
        note("w2_prev", rf["w2"][:-1] != gb["w2_prev"])
        note("use_all", rp["r"]["use_all"] != gb["use_all"])
        note("nbins", rp["r"]["nbins"] != gb["nbins"])
        note("hist_hash", rp["r"]["hist_hash"] != gb["hist_hash"])
        note("nticket", rp["r"]["nticket"] != gb["nticket"])
        # ticket COUNTS are order-independent: must match everywhere
        k = np.arange(3)[None, None, :]
        live = k < gb["nticket"][:, :, None]
        note("ticket_counts", (rp["r"]["tcnt"] != gb["ticket"]["cnt"][:, :, :3]) & live)
        # ticket OFFSETS must match wherever that entry's count is shared with no other bin (nge == ngt + 1)
        unique = live & (gb["nge"][:, :, :3] == gb["ngt"][:, :, :3] + 1)
        off_diff = (rp["r"]["tdx"] != gb["ticket"]["dx"][:, :, :3]) | (rp["r"]["tdy"] != gb["ticket"]["dy"][:, :, :3])
        note("ticket_offsets_where_unique", off_diff & unique)
    active = (gb["n_curr"] > 0).sum(axis=1) if n > 1 else np.zeros(0, np.int64)
    note("active", rp["active"] != active)
    gvalid = (go["flags"] & RB_OFFSET_VALID) != 0
    flagged = (go["flags"] & RB_OFFSET_TIE_SENSITIVE) != 0
    rvalid = rp["valid"] != 0
    differ = (rvalid != gvalid) | (rvalid & gvalid & ((rp["dx"] != go["dx"]) | (rp["dy"] != go["dy"])))
    note("declared_offset_unflagged", differ & ~flagged)
    out = dict(frames_compared=int(n), pairs_compared=int(n - 1), flagged=int(flagged.sum()),
               flagged_and_different=int((differ & flagged).sum()), reference_nullopt=int((~rvalid).sum()),
               ticket_order_differs=int(((off_diff & live).any(axis=(1, 2))).sum()) if n > 1 else 0)
    if check_positions and ref.get("positions") is not None and n > 1:
        # the unmodified frc::collector's (fragment, x, y) against the accumulation of the REFERENCE's declared offsets
        # (pins the digest's pair records to the collector) and of OURS (equal unless a flagged pair differs)
        from remap_b200 import shard
        roff = np.zeros(n - 1, go.dtype)
        roff["dx"], roff["dy"], roff["flags"] = rp["dx"], rp["dy"], rvalid.astype(np.uint32)
        note("collector_positions_vs_reference_offsets", (shard.positions(roff) != ref["positions"]).any(axis=1))
        if out["flagged_and_different"] == 0:
            note("collector_positions", (shard.positions(go) != ref["positions"]).any(axis=1))
    out["mismatches"] = int(sum(v["count"] for v in bad.values()))
    out["mismatch_detail"] = bad
    return out


# ---- numpy mirror of the digest definitions (pins them to the C restatement on the CPU) ----------------------
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix(z):
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def median_hash(median: np.ndarray) -> int:
    flat = median.reshape(-1).astype(np.uint64)
    idx = np.nonzero(flat)[0].astype(np.uint64)
    with np.errstate(over="ignore"):
        return int(splitmix((idx << np.uint64(8)) | flat[idx]).sum(dtype=np.uint64))


def kp_hash(kps: np.ndarray, nreg: int = 8):
    """kps: oracle KP_DTYPE records (code, weight, x, y, region_mask) -> (kp_hash, insertions)"""
    if len(kps) == 0:
        return 0, 0
    code = np.zeros((len(kps), 16), np.uint8)
    code[:, :13] = kps["code"]
    lo = code[:, :8].copy().view("<u8").reshape(-1)
    hi = code[:, 8:16].copy().view("<u8").reshape(-1)
    ch = splitmix(lo ^ splitmix(hi))
    total, ins = np.uint64(0), 0
    with np.errstate(over="ignore"):
        for r in range(nreg):
            sel = ((kps["region_mask"] >> r) & 1) == 1
            if not sel.any():
                continue
            a = kps["x"][sel].astype(np.uint64) | (kps["y"][sel].astype(np.uint64) << np.uint64(16)) | (np.uint64(r) << np.uint64(32))
            total = total + splitmix(a ^ ch[sel]).sum(dtype=np.uint64)
            ins += int(sel.sum())
    return int(total), ins
