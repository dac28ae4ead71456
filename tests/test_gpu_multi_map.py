"""GPU, world_size 2: multi-GPU map assembly and pass 2 on frame shards.  Every rank registers
its contiguous frame range, the pair results are gathered to rank 0, positions go back to everyone, every rank
blits ITS frames into a partial dot map of the fragment, one reduction sums the partial maps on rank 0, blend
there.  Must equal one process on the whole sequence (numpy / C restatement).  Two GPUs: NCCL; one GPU (the
driver's test box): both ranks share cuda:0 and the reduction goes through gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, backend, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import remap_b200
    from remap_b200 import PLACEMENT_DTYPE, shard, synth
    ngpu = torch.cuda.device_count()
    dev = rank % ngpu
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        W, H = 160, 112
        seq = synth.scrolling_tilemap(n_frames, W, H, seed=71, sprites=4, world_w=400, world_h=304)
        first, end, p0, p1 = shard.shard_range(n_frames, world, rank)
        with remap_b200.Registrar(W, H, max_frames=end - first, device=dev) as reg:
            reg.upload(seq.frames[first:end])
            off, _ = reg.register(end - first)
            local = torch.from_numpy(np.ascontiguousarray(off).view(np.int32).reshape(-1, 3).copy())
            allo = shard.gather_offsets(local.to(f"cuda:{dev}") if backend == "nccl" else local, n_frames)
            box = [None]
            if rank == 0:
                pos = shard.positions(allo)
                assert (pos[:, 0] == 0).all(), "one fragment expected"
                box[0] = (pos, shard.fragment_extents(pos[:, 1:], W, H))
            dist.broadcast_object_list(box, src=0)
            pos, (zx, zy, mw, mh) = box[0]
            lo = rank * n_frames // world                      # own frames [lo, end): the overlap frame is the predecessor's
            own = np.arange(lo, end)
            pl = np.zeros(len(own), PLACEMENT_DTYPE)
            pl["frame"], pl["x"], pl["y"] = own - first, pos[own, 1] - zx, pos[own, 2] - zy
            reg.blit_blend(pl, mw, mh, want_dots=False)
            handles = shard.exchange_map_handles(reg)
            fused = shard.reduce_fragment_map_fused(reg, handles=handles)  # one kernel over the peers' maps (CUDA IPC)
            # a pass-2 call with few frames, then one with more (its foreground store grows), between two fused
            # reductions that reuse the cached handles: the peers' map mappings must survive the growth
            half = max(len(pl) // 2, 1)
            reg.filter_fragment(pl[:half], mw, mh, background=np.zeros((mh, mw), np.uint8), want_dots=False)
            shard.reduce_fragment_map_fused(reg, handles=handles)
            reg.filter_fragment(pl, mw, mh, background=np.zeros((mh, mw), np.uint8), want_dots=False)
            shard.reduce_fragment_map_fused(reg, handles=handles)
            reg.blit_blend(pl, mw, mh, want_dots=False)
            again = shard.reduce_fragment_map_fused(reg, handles=handles)
            if rank == 0:
                for a, b in zip(fused, again):
                    assert np.array_equal(a, b), "fused reduction after the foreground store grew differs"
            reg.blit_blend(pl, mw, mh, want_dots=False)          # the partial map again: rank 0's now holds the sum
            plain = shard.reduce_fragment_map(reg)
            if rank == 0:
                for a, b in zip(fused, plain):
                    assert np.array_equal(a, b), "fused peer-memory reduction differs from the NCCL / gloo one"
            bg = [plain[1] if rank == 0 else None]
            dist.broadcast_object_list(bg, src=0)                # the background of pass 2 is the blend of ALL frames
            reg.filter_fragment(pl, mw, mh, background=bg[0], want_dots=False)
            filt = shard.reduce_fragment_map(reg)
            if rank == 0:
                q.put((pos, (zx, zy, mw, mh), plain, filt))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_multi_rank_map_assembly_equals_single_process(world):
    """world 2: the single-destination peer kernel; world 3: reduce-scatter over all links, then gather."""
    from oracle import oracle
    from remap_b200 import synth
    n_frames = 41
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    pos, (zx, zy, mw, mh), plain, filt = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    seq = synth.scrolling_tilemap(n_frames, 160, 112, seed=71, sprites=4, world_w=400, world_h=304)
    want = oracle.assemble_fragment(seq.frames, pos[:, 1:])
    assert want["zero"] == (zx, zy) and want["dots"].shape[:2] == (mh, mw)
    assert np.array_equal(plain[0], want["dots"]) and np.array_equal(plain[1], want["image"]) and np.array_equal(plain[2], want["mask"])
    cfg = oracle.config(160, 112)
    med = np.stack([oracle.extract(cfg, f)[0] for f in seq.frames])
    wf = oracle.filter_fragment(seq.frames, med, pos[:, 1:] - np.array([zx, zy]), mw, mh)
    assert np.array_equal(wf["background"], plain[1])
    assert np.array_equal(filt[0], wf["dots"])
