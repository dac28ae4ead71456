"""Multi-GPU sharding of the registration path.

kpe is independent per frame and kpm per consecutive pair; the only sequential dependency of the
reference's loop is ``position_ += off`` with a reset when no offset was declared
(src/frc.hpp:109-115,124-127).  So a sequence of N frames is cut into contiguous frame ranges, one
per rank; every rank but the first also takes the last frame of its predecessor (one-frame overlap)
so that every consecutive pair belongs to exactly one rank.  There is NO data-path collective while
registering; the only exchange is one gather of the 12-byte pair results to rank 0 (NCCL over
NVLink on GPUs, gloo in the CPU tests), followed by a segmented scan there that turns offsets into
(fragment, x, y) positions for map assembly (mpb).
"""
from __future__ import annotations

import numpy as np

from .api import OFFSET_DTYPE
from ._lib import RB_OFFSET_VALID


def shard_range(n_frames: int, world: int, rank: int):
    """-> (first_frame, end_frame, first_pair, end_pair): this rank registers frames
    [first_frame, end_frame) and owns pairs [first_pair, end_pair) of the global sequence
    (pair i = frames (i, i+1))."""
    lo = rank * n_frames // world
    hi = (rank + 1) * n_frames // world
    first = lo - 1 if lo > 0 else 0   # ranks whose predecessors are all empty start the sequence
    if hi <= lo:
        return lo, lo, max(lo - 1, 0), max(lo - 1, 0)
    return first, hi, first, hi - 1


def gather_offsets(local, n_frames: int, group=None, device=None):
    """Gathers every rank's pair results to rank 0 (one collective per call).

    local: (pairs_r,) OFFSET_DTYPE numpy array, or an int32 torch tensor of shape (pairs_r, 3)
    living on `device` (e.g. a zero-copy view of rb_offsets_device).  Returns on rank 0 the
    (n_frames - 1,) OFFSET_DTYPE array of the whole sequence, None elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_frames, world, r) for r in range(world)]
    counts = [s[3] - s[2] for s in sizes]
    cap = max(max(counts), 1)
    if isinstance(local, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(local).view(np.int32).reshape(-1, 3))
        if device is not None:
            t = t.to(device)
    else:
        t = local
    assert t.shape[0] == counts[rank], (t.shape, counts[rank])
    buf = torch.zeros((cap, 3), dtype=torch.int32, device=t.device)
    buf[:t.shape[0]] = t
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    parts = [o[:c].cpu().numpy() for o, c in zip(out, counts)]
    allp = np.concatenate(parts, 0) if parts else np.zeros((0, 3), np.int32)
    return np.ascontiguousarray(allp).view(OFFSET_DTYPE).reshape(-1)


def positions(offsets) -> np.ndarray:
    """The reference's accumulation (src/frc.hpp:109-115): (N, 3) int32 [fragment, x, y]; a pair
    without a declared offset starts a new fragment at (0, 0)."""
    valid = (offsets["flags"] & RB_OFFSET_VALID) != 0
    n = len(offsets) + 1
    frag = np.zeros(n, np.int64)
    frag[1:] = np.cumsum(~valid)
    dx = np.where(valid, offsets["dx"], 0).astype(np.int64)
    dy = np.where(valid, offsets["dy"], 0).astype(np.int64)
    cx = np.concatenate([[0], np.cumsum(dx)])
    cy = np.concatenate([[0], np.cumsum(dy)])
    # subtract the running sum at the start of each fragment
    start = np.zeros(n, np.int64)
    start[1:] = np.where(~valid, np.arange(1, n), 0)
    start = np.maximum.accumulate(start)
    out = np.stack([frag, cx - cx[start], cy - cy[start]], 1).astype(np.int32)
    return out
