"""Generate tests/golden/*.npz by running the REFERENCE's own compiled code.

Run in the dev container (needs /root/reference, via `make -C oracle ref`):

    python tests/golden/make_golden.py

Every array named ``out_*`` / ``wire_*`` in the fixtures was produced by a function
of the reference's unmodified translation units (oracle/_ref/libpcs_ref_*.so):
copyPointCloudXYZRGBToBufferSIMD / sendXYZRGBPointcloud
(src/pcs-camera-optimized.cpp:363,669), sendStitchToUnity
(src/pcs-multicamera-client.cpp:373), convertBufferToPointCloudXYZRGB /
convertPointCloudXYZRGBToBuffer / updateCloudXYZRGB / send_stitchedXYZRGB
(src/pcs-multicamera-optimized.cpp:226,251,268,299).  Inputs are seeded.
The fixtures pin oracle/pcs_oracle.c (tests/test_oracle.py) and, on the GPU box
where /root/reference does not exist, the CUDA kernels (tests/test_*_gpu.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402
from pointcloud_stitching_b200 import synth  # noqa: E402

oracle.build(ref=True)
RC, RCL, RO, R = oracle.ref_camera(), oracle.ref_client(), oracle.ref_optimized(), oracle.restatement()
assert RC and RCL and RO, "reference libraries missing: run `make -C oracle ref`"


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-28s %7.1f KB" % (name, os.path.getsize(path) / 1024))


# ---- camera side: vertices/texcoords that came out of the deprojection spec ----
for name, (w, h), trans, tf, cutoff in [
    ("pack_96x64_identity", (96, 64), (0, 0, 0), synth.TF_CAMERA, False),
    ("pack_96x64_baseline", (96, 64), synth.D2C_BASELINE, synth.TF_STITCH[3], False),
    ("pack_96x64_cutoff", (96, 64), synth.D2C_BASELINE, synth.TF_CAMERA, True),
]:
    cal = oracle.make_calib(w, h, translation=trans)
    z = synth.depth_frame(w, h, cam=7, frame=len(name), lo=300, hi=3000 if cutoff else 6000)
    col = synth.color_frame(w, h, cam=7, frame=len(name))
    xyz, uv = R.deproject(cal, z)
    out = RC.pack(xyz, uv, col, w, h, 3, w * 3, tf, cutoff=cutoff, threads=1)
    save(name, w=w, h=h, translation=np.float32(trans), z16=z, color=col, xyz=xyz, uv=uv, tf=tf,
         cutoff=cutoff, out_records=out)

# ---- camera side: adversarial vertices / texcoords (no deprojection involved) ----
rng = np.random.default_rng(1234)
n, w, h, stride = 2048, 40, 24, 40 * 3 + 8
xyz = rng.uniform(-8, 8, (n, 3)).astype(np.float32)
uv = rng.uniform(-0.25, 1.25, (n, 2)).astype(np.float32)
xyz[:64] = rng.uniform(-40, 40, (64, 3))            # int16 wrap (|v*1000| > 32767)
uv[64:72] = [np.nan, 0.5]
uv[72:80] = [np.inf, -np.inf]
uv[80:88] = [1e30, -1e30]
xyz[88:96] = [np.nan, 1.0, np.inf]
xyz[96:104] = [3e6, -3e6, 2.2e6]                    # *1000 beyond int32: x86 "indefinite"
uv[104:112] = np.float32([[k / w, k / h] for k in range(8)])           # exact pixel centres' lower edge
uv[112:120] = np.float32([[(k + .5) / w, (k + .5) / h] for k in range(8)])
uv[120:128] = np.float32([[np.nextafter(np.float32((k + .5) / w), np.float32(0)), 0.999999] for k in range(8)])
col = rng.integers(0, 256, (h, stride), dtype=np.uint8)
for cut in (False, True):
    out = RC.pack(xyz, uv, col, w, h, 3, stride, synth.TF_STITCH[6], cutoff=cut, threads=1)
    save("pack_adversarial" + ("_cutoff" if cut else ""), w=w, h=h, stride=stride, xyz=xyz, uv=uv,
         color=col, tf=synth.TF_STITCH[6], cutoff=cut, out_records=out)

# ---- sendXYZRGBPointcloud: buffer image (memset + offset 4) and the wire bytes ----
w, h = 64, 32
cal = oracle.make_calib(w, h)
z, col = synth.depth_frame(w, h, 3, 1), synth.color_frame(w, h, 3, 1)
xyz, uv = R.deproject(cal, z)
size, buf, _ = RC.send(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA)
size2, buf2, wire = RC.send(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, wire=True)
assert size == size2 == w * h * 10
# keep the head of the 10 MB buffer plus both sides of the 5 000 000-byte memset edge
save("send_64x32", w=w, h=h, z16=z, color=col, xyz=xyz, uv=uv, tf=synth.TF_CAMERA, size=size,
     out_head_nosend=buf[: w * h * 5 + 64], out_memset_edge=buf[2499990:2500010],
     out_head_send=buf2[: w * h * 5 + 64], wire_bytes=wire)

# ---- stitch side ----
rec = RC.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA)          # 2048 records
kw = {}
for d in (1, 2, 3, 4, 7):
    kw["out_raw_d%d" % d] = RCL.raw_stitch_1cam(rec, d)
for d in (1, 2, 4):
    kw["out_unpack_d%d" % d] = RO.unpack(rec, d).view(np.uint8)
    for k in (0, 5):
        kw["out_pcl_d%d_tf%d" % (d, k)] = RO.pcl_stitch_1cam(rec, synth.TF_STITCH[k], d)
        assert np.array_equal(kw["out_pcl_d%d_tf%d" % (d, k)], RCL.pcl_stitch_1cam(rec, synth.TF_STITCH[k], d))
save("stitch_2048", records=rec, **kw)

# ---- F10: the int16 -> /1000.0f -> *1000.0f -> int16 round trip over all values ----
# inputs are reconstructible (see tests/test_oracle.py::roundtrip_inputs); only the
# sparse difference out - in is stored.
allv = np.zeros((65536, 5), np.int16)
allv[:, 0] = np.arange(-32768, 32768)
allv[:, 1] = allv[::-1, 0]
allv[:, 2] = np.roll(allv[:, 0], 12345)
allv[:, 3] = np.arange(65536).astype(np.uint16).view(np.int16)
allv[:, 4] = (np.arange(65536) % 251).astype(np.int16)
rt = RO.repack(RO.unpack(allv))
save("roundtrip_all_int16", out_minus_in=(rt.astype(np.int32) - allv.astype(np.int32)).astype(np.int8))
