#!/usr/bin/env python
"""BASELINE.json config #5: 20 synthetic 848x480 camera streams, voxel-grid merge (10 mm), on N GPUs.

    python tools/bench_config5.py                                   # N = 1
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config5.py

Per step and frame: every rank runs K1 over all 20 cameras (its own frames from HBM, the peers' through
NVLink peer memory -- the pull exchange, multigpu.SymmetricFrameSet), which leaves the whole stitched
cloud on every rank; the voxel merge is then sharded by z-slab (multigpu.sharded_voxel_merge): every
rank sorts and reduces 1/N of the points.  Two figures: slabs left where they are (each GPU holds a
z-range of the grid), and slabs exchanged so that every rank ends with the whole merged cloud.
Prints one JSON line on rank 0 and writes gpurun_out/config5_n<N>.json.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pointcloud_stitching_b200 as pcs  # noqa: E402
from pointcloud_stitching_b200 import multigpu, synth  # noqa: E402

W, H, CAMS, FRAMES, LEAF = 848, 480, 20, 2, 10


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    npts = W * H
    layout = multigpu.StitchLayout([npts] * CAMS, world)
    ctx = pcs.Context(device=local, max_streams=CAMS)
    for cam in range(CAMS):
        ctx.set_stream(cam, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE))
    stitched = [multigpu.StitchedBuffer(layout, rank, dev) for _ in range(FRAMES)]
    if world > 1:
        fset = multigpu.SymmetricFrameSet(layout, rank, dev, W, H, FRAMES, stride=((W * 3 + 15) // 16) * 16)
        stride = fset.stride
        for cam in layout.cams_of[rank]:
            for f in range(FRAMES):
                col = np.zeros((H, stride), np.uint8)
                col[:, :W * 3] = synth.color_frame(W, H, cam, f)
                fset.upload(cam, f, synth.depth_frame(W, H, cam, f), col)
        for cam in range(CAMS):
            d = pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE)
            d.color_stride = stride
            ctx.set_stream(cam, d)
        torch.cuda.synchronize()
        fset.barrier()
        batch = ctx.batch(fset.pull_jobs(stitched))
    else:
        keep, jobs = [], []
        for f in range(FRAMES):
            for cam in range(CAMS):
                z = torch.from_numpy(synth.depth_frame(W, H, cam, f).view(np.int16)).to(dev)
                c = torch.from_numpy(synth.color_frame(W, H, cam, f)).to(dev)
                keep.append((z, c))
                jobs.append((cam, z.data_ptr(), c.data_ptr(), stitched[f].slot_ptr(cam)))
        batch = ctx.batch(jobs)
    n = CAMS * npts
    out = torch.zeros(n * 5, dtype=torch.int16, device=dev)
    cs = torch.cuda.current_stream().cuda_stream
    res = {"config": "20 x 848x480, leaf %d mm, %d frames per step" % (LEAF, FRAMES), "n_gpus": world,
           "points_per_frame": n, "k1_launches": batch.launches}

    def step(gather):
        batch.run(cs)
        if world > 1:
            fset.barrier()
        nv = mine = 0
        for f in range(FRAMES):
            nv, mine = multigpu.sharded_voxel_merge(ctx, stitched[f].payload.data_ptr(), n, LEAF, rank, world, out, cs,
                                                    gather=gather)
        return nv, mine

    for gather in (False, True):
        for _ in range(3):
            nv, mine = step(gather)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            nv, mine = step(gather)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res["gathered" if gather else "sharded"] = {"ms_per_step": ms, "mpoints_s": n * FRAMES / ms / 1e3,
                                                    "voxels": nv, "voxels_this_rank": mine}
    # K1 alone (all 20 cameras on every rank)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        batch.run(cs)
        if world > 1:
            fset.barrier()
    e1.record()
    torch.cuda.synchronize()
    res["k1_only_ms_per_step"] = e0.elapsed_time(e1) / 10
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "config5_n%d.json" % world), "w") as f:
            json.dump(res, f)
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
