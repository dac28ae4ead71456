"""CPU, world_size 2 over gloo: the multi-GPU host logic (remap_b200/shard.py).  Every rank takes its
contiguous frame range with the one-frame overlap, registers it (here with the ORACLE standing in for
the kernels -- this test is about sharding, the gather and the position scan), gathers the pair
results to rank 0, and rank 0 must hold exactly what one process gets on the whole sequence."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _offsets_from_oracle(res):
    from remap_b200 import OFFSET_DTYPE, RB_OFFSET_TIE_SENSITIVE, RB_OFFSET_VALID
    out = np.zeros(len(res), OFFSET_DTYPE)
    out["dx"], out["dy"] = res["dx"], res["dy"]
    out["flags"] = (res["valid"] != 0) * RB_OFFSET_VALID + (res["tie_sensitive"] != 0) * RB_OFFSET_TIE_SENSITIVE
    return out


def _worker(rank, world, port, n_frames, cut_every, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle
    from remap_b200 import shard, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        first, end, p0, p1 = shard.shard_range(n_frames, world, rank)
        seq = synth.scrolling_tilemap(n_frames, 200, 136, seed=4, cut_every=cut_every, frame_range=(first, end))
        res = oracle.register(oracle.config(200, 136), seq.frames)["results"]
        local = _offsets_from_oracle(res)
        assert len(local) == p1 - p0
        allo = shard.gather_offsets(local, n_frames)
        if rank == 0:
            q.put(allo)
        else:
            assert allo is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,cut_every", [(41, 0), (60, 13)])
def test_two_rank_gather_equals_single_process(n_frames, cut_every):
    from oracle import oracle
    from remap_b200 import shard, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, cut_every, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seq = synth.scrolling_tilemap(n_frames, 200, 136, seed=4, cut_every=cut_every)
    want = _offsets_from_oracle(oracle.register(oracle.config(200, 136), seq.frames)["results"])
    assert np.array_equal(got, want)
    pos = shard.positions(got)
    # the reference's accumulation, literally (src/frc.hpp:108-115,124-127)
    frag, x, y = 0, 0, 0
    for i in range(1, n_frames):
        if want["flags"][i - 1] & 1:
            x += int(want["dx"][i - 1]); y += int(want["dy"][i - 1])
        else:
            frag += 1; x = y = 0
        assert tuple(pos[i]) == (frag, x, y)
    if cut_every:
        assert pos[-1, 0] >= 1


def test_shard_ranges_cover_every_pair_once():
    from remap_b200 import shard
    for n in (2, 3, 17, 100, 20000):
        for world in (1, 2, 4, 8):
            pairs = []
            for r in range(world):
                first, end, p0, p1 = shard.shard_range(n, world, r)
                assert p1 - p0 == max(end - first - 1, 0)
                pairs += list(range(p0, p1))
            assert pairs == list(range(n - 1)), (n, world)
