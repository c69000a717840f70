#!/usr/bin/env python
"""K1 at other stream geometries than the headline 1280x720 (secondary measurement): fraction of the
HBM copy peak for the TMA-pipelined kernel and, where it does not apply, the direct kernel.
Writes gpurun_out/k1_geom.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pointcloud_stitching_b200 as pcs  # noqa: E402
from pointcloud_stitching_b200 import synth  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

CASES = [
    ("1280x720 baseline", dict(w=1280, h=720, translation=synth.D2C_BASELINE), 64),
    ("848x480 baseline (config #5)", dict(w=848, h=480, translation=synth.D2C_BASELINE), 160),
    ("848x480 aligned", dict(w=848, h=480), 160),
    ("640x480 baseline", dict(w=640, h=480, translation=synth.D2C_BASELINE), 192),
    ("1920x1080 baseline", dict(w=1920, h=1080, translation=synth.D2C_BASELINE), 32),
    ("1280x720, 0.1 deg rotated extrinsics (windowed pipelined kernel)",
     dict(w=1280, h=720, translation=(0.0148, 0.0002, 0.0003),
          rotation=(0.9999975, 0.0010015, 0.0019985, -0.0009985, 0.9999984, -0.001502, -0.002, 0.0015, 0.9999969)), 64),
    ("1280x720, 1 deg rotated extrinsics (direct kernel: window margin too large)",
     dict(w=1280, h=720, translation=(0.015, -0.002, 0.001),
          rotation=(0.99988, 0.0149, 0.0051, -0.0150, 0.99984, 0.0099, -0.0049, -0.0100, 0.99994)), 64),
    ("1280x720 depth + 1920x1080 colour (windowed pipelined kernel)", dict(w=1280, h=720, cw=1920, ch=1080,
                                                                          translation=synth.D2C_BASELINE), 64),
    # the reference's -c flag (src/pcs-camera-optimized.cpp:499-577): keep flags + order-preserving compaction
    ("1280x720 baseline, -c cutoff (z in (0, 1.5], x in (-2, 2])", dict(w=1280, h=720, translation=synth.D2C_BASELINE, cutoff=True), 64),
    ("1280x720 baseline, RGBA colour (general kernel)", dict(w=1280, h=720, translation=synth.D2C_BASELINE, bpp=4), 64),
]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if any(a in c[0] for a in sys.argv[1:])]


def main():
    res = []
    cs = torch.cuda.current_stream().cuda_stream
    for name, kw, n_frames in CASES:
        kw = dict(kw)
        w, h = kw.pop("w"), kw.pop("h")
        cw, ch = kw.get("cw", w), kw.get("ch", h)
        ctx = pcs.Context(device=0, max_streams=1)
        ctx.set_stream(0, pcs.stream_desc(w, h, tf=synth.TF_STITCH[0], **kw))
        distinct = min(n_frames, 48)      # enough distinct input that a launch streams from HBM
        z = torch.from_numpy(np.stack([synth.depth_frame(w, h, 0, f) for f in range(distinct)]).view(np.int16)).cuda()
        c = torch.from_numpy(np.stack([synth.color_frame(cw, ch, 0, f, bpp=kw.get("bpp", 3)) for f in range(distinct)])).cuda()
        pay = torch.zeros(n_frames * w * h * 5, dtype=torch.int16, device="cuda")
        cnt = torch.zeros(n_frames, dtype=torch.int32, device="cuda")
        jobs = [(0, z[f % distinct].data_ptr(), c[f % distinct].data_ptr(), pay.data_ptr() + f * w * h * 10, None,
                 cnt.data_ptr() + 4 * f) for f in range(n_frames)]
        b = ctx.batch(jobs)
        for _ in range(3):
            b.run(cs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            b.run(cs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        pts = n_frames * w * h
        kept = int(cnt.sum().item())
        alg = pts * 2 + kept * 10 + n_frames * cw * ch * kw.get("bpp", 3)
        res.append({"case": name, "frames_per_launch": n_frames, "launches": b.launches, "ms": ms, "records_kept": kept / pts,
                    "mpoints_s": pts / ms / 1e3, "GBps": alg / ms / 1e6, "frac_of_copy_peak": alg / ms / 1e6 / PEAK})
        print(res[-1])
        b.close()
        ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "k1_geom.json" if len(sys.argv) == 1 else "k1_geom_part.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
