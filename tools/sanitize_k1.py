#!/usr/bin/env python
"""Small K1 runs for compute-sanitizer (memcheck, racecheck): the pipelined kernel in its three forms -- identity
extrinsics, x baseline (exact tap chain, row map with 1.5x colour), rotated calibration (guarded taps: segment windows,
per-stage tables written by the producer warp, exact re-evaluation patching the slab) -- each checked against the oracle.

    compute-sanitizer --tool racecheck python tools/sanitize_k1.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import pointcloud_stitching_b200 as pcs  # noqa: E402
from pointcloud_stitching_b200 import synth  # noqa: E402

R = oracle.restatement()
ctx = pcs.Context(device=0, max_streams=4, kernel_variant=2)
cs = torch.cuda.current_stream().cuda_stream
cases = [dict(w=256, h=24), dict(w=256, h=24, translation=synth.D2C_BASELINE),
         dict(w=256, h=24, cw=384, ch=36, translation=synth.D2C_BASELINE),
         dict(w=384, h=24, cw=512, ch=32, translation=(0.015, 0.0003, -0.0002), rotation=synth.rotation_colmajor(0.004, -0.003, 0.005))]
n_checked = 0
for s, kw in enumerate(cases):
    kw = dict(kw)
    w, h = kw.pop("w"), kw.pop("h")
    cw, ch = kw.pop("cw", w), kw.pop("ch", h)
    cal = oracle.make_calib(w, h, cw, ch, **kw)
    ctx.set_stream(s, pcs.stream_desc(w, h, cw, ch, tf=synth.TF_STITCH[s], **kw))
    keep, jobs = [], []
    for f in range(3):
        z = synth.depth_frame(w, h, s, f, lo=1 if f == 0 else 300)
        col = synth.color_frame(cw, ch, s, f)
        dz, dc = torch.from_numpy(z.view(np.int16)).cuda(), torch.from_numpy(col).cuda()
        pay = torch.zeros(w * h * 5, dtype=torch.int16, device="cuda")
        keep.append((z, col, dz, dc, pay))
        jobs.append((s, dz.data_ptr(), dc.data_ptr(), pay.data_ptr()))
    b = ctx.batch(jobs)
    b.run(cs)
    torch.cuda.synchronize()
    for z, col, _, _, pay in keep:
        assert np.array_equal(pay.cpu().numpy().reshape(-1, 5), R.frame(cal, z, col, 3, cw * 3, synth.TF_STITCH[s])), s
        n_checked += 1
    b.close()
print("sanitize_k1: OK", n_checked, "frames")
