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


def gather_offsets_device(local, n_frames: int, group=None, out=None):
    """The collective alone, asynchronous on the current stream: gathers every rank's (pairs_r, 3) int32
    device tensor to rank 0.  Returns (list of per-rank padded tensors on rank 0 | None, counts);
    `out` lets the caller reuse the receive buffers between steps."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [s[3] - s[2] for s in (shard_range(n_frames, world, r) for r in range(world))]
    cap = max(max(counts), 1)
    assert local.shape[0] == counts[rank], (local.shape, counts[rank])
    if local.shape[0] == cap:
        buf = local
    else:
        buf = torch.zeros((cap, 3), dtype=torch.int32, device=local.device)
        buf[:local.shape[0]] = local
    if rank == 0 and out is None:
        out = [torch.empty((cap, 3), dtype=torch.int32, device=local.device) for _ in range(world)]
    dist.gather(buf, out if rank == 0 else None, dst=0, group=group)
    return (out if rank == 0 else None), counts


def gather_offsets(local, n_frames: int, group=None, device=None):
    """Gathers every rank's pair results to rank 0 (one collective per call).

    local: (pairs_r,) OFFSET_DTYPE numpy array, or an int32 torch tensor of shape (pairs_r, 3)
    living on `device` (e.g. a zero-copy view of rb_offsets_device).  Returns on rank 0 the
    (n_frames - 1,) OFFSET_DTYPE array of the whole sequence, None elsewhere.
    """
    import torch

    if isinstance(local, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(local).view(np.int32).reshape(-1, 3))
        if device is not None:
            t = t.to(device)
    else:
        t = local
    out, counts = gather_offsets_device(t, n_frames, group)
    if out is None:
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


def fragment_extents(pos, width: int, height: int):
    """The dot-map geometry fgm::fragment ends up with after blitting frames at `pos` in order
    (src/fgm.hpp:190-233: ensure/extend grow the map in whole frame-sized steps, `zero_` moves with it).

    pos: (n, 2) int positions of ONE fragment (as frc::collector accumulates them, first frame at 0, 0).
    -> (zero_x, zero_y, map_w, map_h); a frame at pos p lies at p - zero inside the map."""
    zero = [0, 0]
    dim = [width, height]  # fragment(step) starts with a map of one step (src/fgm.hpp:44-47)
    step = (width, height)
    for p in np.asarray(pos, np.int64):
        for k in (0, 1):
            lo = hi = 0
            if p[k] < zero[k]:  # src/fgm.hpp:207-211
                lo = _round_up_step(zero[k] - int(p[k]), step[k])
            required = int(p[k]) + step[k]  # src/fgm.hpp:213-221
            if required > 0:
                limit = zero[k] + dim[k]
                if required > limit:
                    hi = _round_up_step(required - limit, step[k])
            zero[k] -= lo  # src/fgm.hpp:223
            dim[k] += lo + hi  # matrix::extend, src/mrl.hpp:131-147
    return zero[0], zero[1], dim[0], dim[1]


def _round_up_step(change: int, step: int) -> int:
    """fgm::fragment::get_step (src/fgm.hpp:228-233)"""
    rest = change % step
    return (change - rest) + (step if rest else 0)


class _DeviceArray:
    """A raw device allocation as something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def reduce_fragment_map(reg, group=None, dst: int = 0, want_dots: bool = True):
    """Multi-GPU map assembly.  Every rank has called ``reg.blit_blend`` (or
    ``reg.filter_fragment``) with ITS frames of one fragment and the fragment's full map geometry; this sums the
    partial dot maps to rank ``dst`` with one reduction (NCCL over NVLink on GPUs; through host memory when the
    process group is gloo, as in the single-GPU tests) and runs ``fgm::fragment::blend`` there.

    The reference's counters are uint16 and wrap (src/fgm.hpp:12-14,94): the partial maps wrap the same way,
    so they are summed as int32 and only the low 16 bits are kept -- equal to one wrapping counter over all frames.
    -> (dots, image, mask) on rank ``dst``, None elsewhere."""
    import torch
    import torch.distributed as dist

    md = reg.map_device()
    count = md["width"] * md["height"] * 16
    dev = torch.device("cuda", torch.cuda.current_device())
    part16 = torch.as_tensor(_DeviceArray(md["dots"], count, "<i2"), device=dev)   # view, not a copy
    wide = part16.to(torch.int32) & 0xFFFF
    if dist.get_backend(group) == "nccl":
        dist.reduce(wide, dst=dst, op=dist.ReduceOp.SUM, group=group)
    else:
        host = wide.cpu()
        dist.reduce(host, dst=dst, op=dist.ReduceOp.SUM, group=group)
        wide = host.to(dev)
    if dist.get_rank(group) != dst:
        return None
    low = wide & 0xFFFF
    part16.copy_((((low + 0x8000) & 0xFFFF) - 0x8000).to(torch.int16))  # same bits as the uint16 value
    torch.cuda.current_stream().synchronize()
    return reg.blend_map(want_dots=want_dots)


def exchange_map_handles(reg, group=None):
    """Every rank's map_export() on every rank.  A handle stays valid until that rank's map scratch grows, so one
    exchange serves all fragments that fit the first one's scratch."""
    import torch.distributed as dist

    handles = [None] * dist.get_world_size(group)
    dist.all_gather_object(handles, reg.map_export(), group=group)
    return handles


def reduce_fragment_map_fused(reg, group=None, dst: int = 0, want_dots: bool = True, handles=None):
    """reduce_fragment_map as ONE kernel over peer memory: the ranks exchange 80-byte CUDA-IPC handles of their
    partial dot maps (all_gather_object), rank ``dst`` reads the peers' maps in place over NVLink, sums and blends
    in a single pass (rb_blend_map_peers); a barrier keeps the peers' maps alive until it is done.  With more than two
    ranks the sum is spread over all links first: every rank reduces one slice of the map over its peers
    (rb_sum_map_slice), then the destination pulls the reduced slices and blends (rb_blend_map_slices).  Pass
    `handles` from an earlier exchange_map_handles() to skip the exchange.
    -> (dots, image, mask) on rank ``dst``, None elsewhere."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if handles is None:
        handles = exchange_map_handles(reg, group)
    # every rank's blit (blit_blend / filter_fragment waits for its own stream) must have finished before a peer
    # reads its map over IPC; with cached handles nothing else orders that
    dist.barrier(group=group)
    out = None
    if world <= 2:  # one kernel on the destination rank: the only peer's map crosses the link once
        if rank == dst:
            out = reg.blend_map_peers([h for r, h in enumerate(handles) if r != dst], want_dots=want_dots)
    else:           # all links at once: every rank reduces one slice, the destination gathers the slices and blends
        reg.sum_map_slice(handles, rank)
        dist.barrier(group=group)
        if rank == dst:
            out = reg.blend_map_slices(handles, rank, want_dots=want_dots)
    dist.barrier(group=group)
    return out
