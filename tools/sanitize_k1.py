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
# the reference's -c in one pass (k1_direct<CUTOFF>: persistent blocks, ticketed tiles, look-back over status words, compaction
# in shared memory): several frames in one launch so that blocks hold two tiles at a time, wide and narrow cutoff boxes
c0 = pcs.Context(device=0, max_streams=2, kernel_variant=0)
for s, box in enumerate([dict(), dict(z_hi=4.0, x_lo=-0.3, x_hi=0.9)]):
    w, h = 512, 96
    cal = oracle.make_calib(w, h, translation=synth.D2C_BASELINE)
    d = pcs.stream_desc(w, h, tf=synth.TF_STITCH[s], translation=synth.D2C_BASELINE, cutoff=True)
    for k, v in box.items():
        setattr(d, k, v)
    c0.set_stream(s, d)
    keep, jobs = [], []
    for f in range(4):
        z = synth.depth_frame(w, h, 10 + s, f)
        col = synth.color_frame(w, h, 10 + s, f)
        dz, dc = torch.from_numpy(z.view(np.int16)).cuda(), torch.from_numpy(col).cuda()
        pay = torch.zeros(w * h * 5, dtype=torch.int16, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        keep.append((z, col, dz, dc, pay, cnt))
        jobs.append((s, dz.data_ptr(), dc.data_ptr(), pay.data_ptr(), None, cnt.data_ptr()))
    b = c0.batch(jobs)
    b.run(cs)
    torch.cuda.synchronize()
    for z, col, _, _, pay, cnt in keep:
        xyz, uv = R.deproject(cal, z)
        want = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_STITCH[s], True) if not box else None
        n = int(cnt.item())
        if want is not None:
            assert n == len(want) and np.array_equal(pay[: n * 5].cpu().numpy().reshape(-1, 5), want), s
        else:
            assert 0 < n < w * h
        n_checked += 1
    b.close()
print("sanitize_k1: OK", n_checked, "frames")
