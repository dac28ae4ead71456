#!/usr/bin/env python3
"""A BASELINE.json config at its STATED size and GPU count, with full-sequence parity against the real reference.

  python tools/run_config.py --config 4                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/run_config.py --config 5 [--frames-total 1000000] [--batch 25000] [--no-parity]

Every rank takes a contiguous frame range of ONE global sequence (one-frame overlap with its predecessor) and walks
it in batches: render the batch on the host (seeded generator, remap_b200/synth.py), then
  e2e     rb_register_host_async from the pinned batch + rb_fetch_offsets      (wall clock, copies inside)
  device  rb_register_async over the now resident batch                        (CUDA events inside the library)
  parity  rb_frame_digests / rb_fetch_ballots / offsets against oracle/_ref/ref_harness digest run on this rank's share
          of the host cores over the very same frames: EVERY frame and pair (tests/digest_check.py)
The 12-byte pair results are gathered to rank 0 (NCCL), which accumulates positions / fragments like frc::collector.
Config 3 adds pass 2 on the resident frames: per-rank partial maps (rb_blit_blend), one reduction, the blended
background back to every rank, rb_filter_fragment per rank, one reduction, blend (src/fdf.hpp:40-75).

TEST / MEASUREMENT TOOL: uses oracle/ as the checker.  Writes gpurun_out/r2_config<K>_n<N>.json on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FULL = {  # total frames of the stated config
    2: 20000, 3: 100000, 4: 5000, 5: 1000000,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[2, 3, 4, 5])
    ap.add_argument("--frames-total", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0, help="frames per batch and rank (default 25,000; 2,500 at 640x480)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-pass2", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import bench
    import digest_check
    import remap_b200
    from remap_b200 import PLACEMENT_DTYPE, RB_OFFSET_TIE_SENSITIVE, RB_OFFSET_VALID, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=900))

    cfgd = bench.CONFIGS[args.config]
    W, H, gen = cfgd["width"], cfgd["height"], dict(cfgd["gen"])
    total = args.frames_total or FULL[args.config]
    first, end, p0, p1 = shard.shard_range(total, world, rank)
    n_rank = end - first
    batch = args.batch or (2500 if W * H > 200000 else 25000)
    if args.config == 3 and not args.no_pass2:
        batch = n_rank  # pass 2 works on the rank's resident frames
    batch = min(batch, n_rank)
    threads = max(1, (os.cpu_count() or 1) // max(local_world, 1))
    t_plan = time.perf_counter()
    plan = synth.TilemapPlan(total, W, H, **gen)
    t_plan = time.perf_counter() - t_plan

    pinned = torch.empty((batch, H, W), dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    offs = np.zeros(n_rank - 1, remap_b200.OFFSET_DTYPE)
    stats = dict(frames_compared=0, pairs_compared=0, mismatches=0, flagged=0, flagged_and_different=0, reference_nullopt=0,
                 detail={})
    t_e2e = t_dev = t_gen = t_ref = 0.0
    kt_sum = {}
    kp_total = 0
    deferred = 0
    lanes = None
    reg = remap_b200.Registrar(W, H, max_frames=batch, device=local_rank, profile=True)
    try:
        at = 0  # frames of this rank done; batch b covers rank frames [at - 1 (overlap), at + m)
        while at < n_rank:
            lo = at - 1 if at > 0 else 0
            m = min(batch, n_rank - lo)
            t0 = time.perf_counter()
            seq = plan.render(first + lo, first + lo + m, out=host[:m])
            t_gen += time.perf_counter() - t0
            for rep in range(2 if at == 0 else 1):  # the very first call also pays allocations: repeat it
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                reg.register_host_async(host[:m])
                o = reg.fetch_offsets(m - 1)
                dt = time.perf_counter() - t0
            t_e2e += dt
            lanes = reg.host_lane_stats
            reg.register_async(m)
            k = reg.kernel_times()
            t_dev += (k["kpe_ms"] + k["kpm_ms"] + k["declare_ms"]) * 1e-3
            for key, v in k.items():
                kt_sum[key] = kt_sum.get(key, 0.0) + v
            o2 = reg.fetch_offsets(m - 1)
            assert np.array_equal(o, o2), "host path and resident path disagree"
            deferred += reg.deferred_count
            kp_total += reg.count_keypoints(m - (1 if at > 0 else 0), first=(1 if at > 0 else 0))
            offs[lo:lo + m - 1] = o
            if not args.no_parity:
                t0 = time.perf_counter()
                ref = digest_check.ref_digest(seq.frames, threads=threads, collector=False)
                t_ref += time.perf_counter() - t0
                res = digest_check.compare(ref, digest_check.gpu_digest(reg, m, offsets=o), check_positions=False)
                own_f = m - (1 if at > 0 else 0)
                stats["frames_compared"] += own_f
                stats["pairs_compared"] += res["pairs_compared"]
                for key in ("mismatches", "flagged", "flagged_and_different", "reference_nullopt"):
                    stats[key] += res[key]
                if res["mismatch_detail"]:
                    stats["detail"][f"rank{rank}_frame{first + lo}"] = res["mismatch_detail"]
            at = lo + m

        # ---- the one exchange of the path: pair results to rank 0, positions there ---------------------------------
        if world > 1:
            allo = shard.gather_offsets(offs, total, device=dev)
        else:
            allo = offs
        pass2 = None
        if args.config == 3 and not args.no_pass2:
            box = [None]
            if rank == 0:
                pos = shard.positions(allo)
                box[0] = pos
            if world > 1:
                dist.broadcast_object_list(box, src=0)
            pos = box[0]
            frag0 = pos[:, 0] == pos[0, 0]
            idx_all = np.nonzero(frag0)[0]
            zx, zy, mw, mh = shard.fragment_extents(pos[idx_all, 1:], W, H)
            own_lo = rank * total // world
            own = np.arange(own_lo, end)
            own = own[frag0[own]]
            pl = np.zeros(len(own), PLACEMENT_DTYPE)
            pl["frame"], pl["x"], pl["y"] = own - first, pos[own, 1] - zx, pos[own, 2] - zy
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            bb = reg.blit_blend(pl, mw, mh, want_dots=False)
            if world > 1:
                bgres = shard.reduce_fragment_map(reg, want_dots=False)
                bgbox = [bgres[1] if rank == 0 else None]
                dist.broadcast_object_list(bgbox, src=0)
                bg = bgbox[0]
            else:
                bg = bb[1]
            t_bg = time.perf_counter() - t0
            t0 = time.perf_counter()
            fr = reg.filter_fragment(pl, mw, mh, background=bg, want_dots=False)
            if world > 1:
                out = shard.reduce_fragment_map(reg, want_dots=False)
            else:
                out = (None, fr["image"], fr["mask"])
            t_f = time.perf_counter() - t0
            if rank == 0:
                # the static world seen through the camera: with the sprites filtered out the map must show the world itself
                wx0, wy0 = int(plan.path[0, 0]) + zx, int(plan.path[0, 1]) + zy      # world coordinates of map pixel (0, 0)
                world_img = plan.worlds[0]
                mx0, my0 = max(0, -wx0), max(0, -wy0)                               # the part of the map that lies over the world
                mx1, my1 = min(mw, world_img.shape[1] - wx0), min(mh, world_img.shape[0] - wy0)
                img, msk = out[1][my0:my1, mx0:mx1], out[2][my0:my1, mx0:mx1]
                wv = world_img[wy0 + my0:wy0 + my1, wx0 + mx0:wx0 + mx1]
                bgv = bg[my0:my1, mx0:mx1]
                cov = msk != 0
                pass2 = dict(frames=int(len(idx_all)), map=[int(mw), int(mh)], background_s=t_bg, filter_s=t_f,
                             frames_per_s=float(len(idx_all) / (t_bg + t_f)),
                             kernel_ms=fr["times_ms"], covered_pixels=int(cov.sum()),
                             background_equals_world=float((bgv[cov] == wv[cov]).mean()),
                             filtered_map_equals_world=float((img[cov] == wv[cov]).mean()))
    finally:
        reg.close()

    # ---- reduce the timings / counts -----------------------------------------------------------------------------
    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    dev_s, e2e_s = allmax(t_dev), allmax(t_e2e)
    for key in ("frames_compared", "pairs_compared", "mismatches", "flagged", "flagged_and_different", "reference_nullopt"):
        stats[key] = int(allsum(stats[key]))
    kp_all, deferred_all = allsum(kp_total), int(allsum(deferred))
    details = [None] * world
    if world > 1:
        dist.all_gather_object(details, stats["detail"])
    else:
        details = [stats["detail"]]
    if rank == 0:
        peak, peak_src = bench.measured_peaks()
        valid = (allo["flags"] & RB_OFFSET_VALID) != 0
        pos = shard.positions(allo)
        kpf = kp_all / total
        b_path = 2 * W * H + 60 * kpf + 12
        line = dict(
            config=args.config, workload=cfgd["what"], n_gpus=world, frames_total=total, frames_per_rank=n_rank, batch=batch,
            width=W, height=H,
            value=total / dev_s, unit="frames/s", device_seconds=dev_s,
            e2e=dict(value=total / e2e_s, seconds=e2e_s, lanes_rank0=lanes),
            roofline=dict(frac_path=b_path * total / dev_s / 1e9 / (peak * world),  # per GPU: N GPUs move N x the bytes
                           bytes_per_frame=b_path, keypoints_per_frame=kpf, peak=peak,
                          peak_source=peak_src, kernel_ms_rank0={k: v for k, v in kt_sum.items()}),
            parity=dict(against="oracle/_ref/ref_harness digest (the reference's own kpe::extractor::extract, kpm::match, "
                                "count_offsets, top_offsets) on the same frames",
                        enabled=not args.no_parity, **{k: v for k, v in stats.items() if k != "detail"},
                        pairs_run=total - 1, detail={k: v for d in details for k, v in (d or {}).items()},
                        reference_seconds_rank0=t_ref, reference_threads_per_rank=threads),
            deferred_ballots=deferred_all,
            fragments=int(pos[-1, 0]) + 1, valid_fraction=float(valid.mean()),
            tie_sensitive_fraction=float(((allo["flags"] & RB_OFFSET_TIE_SENSITIVE) != 0).mean()),
            generation_seconds_rank0=t_gen + t_plan, pass2=pass2,
        )
        out = args.out or os.path.join(ROOT, "gpurun_out", f"r2_config{args.config}_n{world}.json")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(line, open(out, "w"), indent=1)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not args.no_parity and stats["mismatches"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
