#!/usr/bin/env python3
"""Multi-GPU map assembly: the cross-rank sum of the partial dot maps + blend, (a) as NCCL reduce between
conversion passes (shard.reduce_fragment_map) and (b) as one kernel over peer memory (reduce_fragment_map_fused).
Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_map_reduce.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import remap_b200  # noqa: E402
from remap_b200 import PLACEMENT_DTYPE, shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, per = 320, 224, 2000
    n = per * world
    seq = synth.scrolling_tilemap(n, W, H, seed=5)
    first, end, _, _ = shard.shard_range(n, world, rank)
    pos = (seq.path - seq.path[0]).astype(np.int64)          # ground-truth positions (the gather is timed elsewhere)
    zx, zy, mw, mh = shard.fragment_extents(pos, W, H)
    lo = rank * n // world
    own = np.arange(lo, end)
    with remap_b200.Registrar(W, H, max_frames=end - first, device=local) as reg:
        reg.upload(seq.frames[first:end])
        pl = np.zeros(len(own), PLACEMENT_DTYPE)
        pl["frame"], pl["x"], pl["y"] = own - first, pos[own, 0] - zx, pos[own, 1] - zy
        res = {}
        reg.blit_blend(pl, mw, mh, want_dots=False)
        handles = shard.exchange_map_handles(reg)
        fused = lambda r, want_dots: shard.reduce_fragment_map_fused(r, want_dots=want_dots, handles=handles)  # noqa: E731
        for name, fn in (("nccl_reduce", shard.reduce_fragment_map), ("fused_peer_kernel", fused)):
            ts, out = [], None
            for r in range(4):
                reg.blit_blend(pl, mw, mh, want_dots=False)      # fresh partial maps
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = fn(reg, want_dots=False)
                dist.barrier()
                ts.append(time.perf_counter() - t0)
            res[name] = round(min(ts[1:]) * 1e3, 3)
            if rank == 0:
                res[name + "_checksum"] = int(out[1].astype(np.int64).sum())
        if rank == 0:
            print(json.dumps(dict(n_gpus=world, frames=n, map=[int(mw), int(mh)], dots_MB_per_rank=round(mw * mh * 32 / 1e6, 1),
                                  ms=res)), file=sys.__stdout__, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
