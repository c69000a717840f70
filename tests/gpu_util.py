"""Helpers shared by the -m gpu parity tests (torch is only used for device memory)."""
import numpy as np
import torch

import oracle
import pointcloud_stitching_b200 as pcs
from pointcloud_stitching_b200 import synth


def dev(a):
    """numpy -> cuda tensor (bytes preserved)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda()
    return torch.from_numpy(a).cuda()


def calib_and_desc(w, h, cw=None, ch=None, tf=synth.TF_CAMERA, translation=(0.0, 0.0, 0.0),
                   rotation=None, bpp=3, stride=None, cutoff=False, **kw):
    cal = oracle.make_calib(w, h, cw, ch, rotation=rotation, translation=translation, **kw)
    desc = pcs.stream_desc(w, h, cw, ch, tf=tf, translation=translation, rotation=rotation, bpp=bpp,
                           stride=stride, cutoff=cutoff, **kw)
    return cal, desc


def small_rotation(rx=0.01, ry=-0.02, rz=0.015):
    """Column-major 3x3 (rs2_extrinsics convention), float32."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx).astype(np.float32)
    return tuple(float(v) for v in R.T.reshape(-1))  # column-major


def run_batch(ctx, jobs_np, variant_streams, float_out=False, want_count=False):
    """jobs_np: list of (stream, z16 ndarray, colour ndarray).  Returns list of
    (records int16[n,5], xyzrgb or None, count or None)."""
    keep, jobs, outs = [], [], []
    for stream, z, col in jobs_np:
        n = z.size
        dz, dc = dev(z), dev(col)
        pay = torch.full((n * 5 + 8,), 0x5A5A, dtype=torch.int16, device="cuda")
        fo = torch.zeros((n, 4), dtype=torch.float32, device="cuda") if float_out else None
        cnt = torch.full((1,), -1, dtype=torch.int32, device="cuda") if want_count else None
        keep.append((dz, dc, pay, fo, cnt))
        jobs.append((stream, dz.data_ptr(), dc.data_ptr(), pay.data_ptr(),
                     fo.data_ptr() if float_out else None, cnt.data_ptr() if want_count else None))
        outs.append((pay, fo, cnt, n))
    b = ctx.batch(jobs)
    b.run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = []
    for pay, fo, cnt, n in outs:
        assert int(pay[n * 5]) == 0x5A5A, "kernel wrote past the payload"
        c = int(cnt.item()) if cnt is not None else n
        res.append((pay[: c * 5].cpu().numpy().reshape(-1, 5), fo.cpu().numpy() if fo is not None else None, c))
    b.close()
    return res
