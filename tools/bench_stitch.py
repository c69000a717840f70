#!/usr/bin/env python
"""Secondary measurements (not the bench.py contract): stitch-side kernels and the voxel
merge on one GPU, CUDA-event timed, inputs resident.  Writes gpurun_out/stitch_bench.json.

    python tools/bench_stitch.py [--cams 4] [--iters 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pointcloud_stitching_b200 as pcs  # noqa: E402
from pointcloud_stitching_b200 import synth  # noqa: E402

W, H = 1280, 720
N = W * H


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cams", type=int, default=16, help="16 x 720p = 147 MB in + 147 MB out: larger than the 126 MB L2")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only-voxel-variant", type=int, default=-1, help="time only this voxel_variant (for ncu runs)")
    ap.add_argument("--skip-stitch", action="store_true", help="voxel merge only")
    ap.add_argument("--scene", default="noise", choices=["noise", "smooth"],
                    help="depth noise: 400 mm (default: ~0.84 voxels per point, the merge's worst case) or 5 mm (a surface)")
    a = ap.parse_args()
    ctx = pcs.Context(device=0, max_streams=a.cams)
    cs = torch.cuda.current_stream().cuda_stream
    # realistic records: run K1 on synthetic frames, one camera each, into separate payloads
    pays, jobs, keep = [], [], []
    for c in range(a.cams):
        ctx.set_stream(c, pcs.stream_desc(W, H, tf=synth.TF_STITCH[c % 8], translation=synth.D2C_BASELINE))
        z = torch.from_numpy(synth.depth_frame(W, H, c, 0, noise_mm=synth.SMOOTH_NOISE_MM if a.scene == "smooth"
                                               else synth.NOISE_MM).view(np.int16)).cuda()
        col = torch.from_numpy(synth.color_frame(W, H, c, 0)).cuda()
        p = torch.zeros(N * 5, dtype=torch.int16, device="cuda")
        keep.append((z, col))
        pays.append(p)
        jobs.append((c, z.data_ptr(), col.data_ptr(), p.data_ptr()))
    b = ctx.batch(jobs)
    b.run(cs)
    torch.cuda.synchronize()
    total = a.cams * N
    st = torch.zeros(16 + total * 10, dtype=torch.uint8, device="cuda")
    cloud = torch.zeros(total * 8, dtype=torch.float32, device="cuda")
    ptrs, ns = [p.data_ptr() for p in pays], [N * 5] * a.cams
    tfs = [synth.TF_STITCH[c % 8] for c in range(a.cams)]
    res = {"cams": a.cams, "points": total, "scene": a.scene}
    if not a.skip_stitch:
        for d in (1, 2, 4):
            ms = timed(lambda: ctx.stitch_raw_dev(ptrs, ns, d, st.data_ptr() + 12, total * 10 + 4, cs), a.iters)
            res["stitch_raw_d%d" % d] = {"ms": ms, "mpoints_s_in": total / ms / 1e3,
                                         "GBps": (total * 10 / d + total * 10 / d) / ms / 1e6}
        ms = timed(lambda: ctx.stitch_pcl_dev(ptrs, ns, 1, tfs, st.data_ptr() + 12, total * 10 + 4, None, cs), a.iters)
        res["stitch_pcl"] = {"ms": ms, "mpoints_s": total / ms / 1e3, "GBps": total * 20 / ms / 1e6}
        ms = timed(lambda: ctx.stitch_pcl_dev(ptrs, ns, 1, tfs, st.data_ptr() + 12, total * 10 + 4, cloud.data_ptr(), cs), a.iters)
        res["stitch_pcl_cloud32"] = {"ms": ms, "mpoints_s": total / ms / 1e3, "GBps": total * 52 / ms / 1e6}
    # voxel merge of the raw-stitched cloud (records at st+16)
    ctx.stitch_raw_dev(ptrs, ns, 1, st.data_ptr() + 12, total * 10 + 4, cs)
    out = torch.zeros(total * 5, dtype=torch.int16, device="cuda")
    nv = [0]

    # voxel_variant: 0 auto (one-sweep sort), 1 (key, index) pair sort, 2 / 3 one-sweep sort with 8- / 10-bit digits,
    # 4 slab partition + bitmap ranking
    for variant, name in ((0, "voxel_merge_10mm"), (1, "voxel_merge_10mm_pair_sort"), (2, "voxel_merge_10mm_sweep8"),
                          (3, "voxel_merge_10mm_sweep10"), (4, "voxel_merge_10mm_msd")):
        if a.only_voxel_variant >= 0 and variant != a.only_voxel_variant:
            continue
        vctx = pcs.Context(device=0, max_streams=1, voxel_variant=variant)

        def vox():
            nv[0] = vctx.voxel_merge_dev(st.data_ptr() + 16, total, 10, out.data_ptr(), cs)
        ms = timed(vox, max(3, a.iters // 4))
        res[name] = {"ms": ms, "mpoints_s_in": total / ms / 1e3, "voxels": nv[0]}
        vctx.close()
    # sharded merge (one z-slab per GPU): what each of 8 ranks would run on the same stitched cloud
    if a.only_voxel_variant < 0:
        vctx = pcs.Context(device=0, max_streams=1)
        ms = timed(lambda: vctx.voxel_slab_plan_dev(st.data_ptr() + 16, total, 10, 8, cs), 5)
        splits, pts = vctx.voxel_slab_plan_dev(st.data_ptr() + 16, total, 10, 8, cs)
        res["voxel_slab_plan_8"] = {"ms": ms, "slab_points": pts}
        slab_ms, slab_nv = [], []
        for r in range(8):
            def one():
                nv[0] = vctx.voxel_merge_slab_dev(st.data_ptr() + 16, total, 10, splits[r], splits[r + 1], out.data_ptr(), cs)
            slab_ms.append(timed(one, 5))
            slab_nv.append(nv[0])
        res["voxel_slab_merge_1of8"] = {"ms_each": slab_ms, "ms_max": max(slab_ms), "voxels": slab_nv,
                                        "mpoints_s_if_8_gpus": total / (ms + max(slab_ms)) / 1e3}
        vctx.close()
    if not a.skip_stitch:
        # K1a: the reference seam itself (vertices + tex coords in), 8 frames of 1280x720 per call
        nf = 8
        xyz = torch.randn(nf * N, 3, device="cuda") * 2.0
        uv = torch.rand(nf * N, 2, device="cuda")
        col = keep[0][1]
        pay1 = torch.zeros(nf * N * 5, dtype=torch.int16, device="cuda")
        ms = timed(lambda: ctx.pack_from_vertices_dev(0, xyz.data_ptr(), uv.data_ptr(), nf * N, col.data_ptr(),
                                                      pay1.data_ptr(), None, cs), a.iters)
        res["k1a_from_vertices"] = {"ms": ms, "mpoints_s": nf * N / ms / 1e3, "GBps": nf * N * 30 / ms / 1e6}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stitch_bench.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
