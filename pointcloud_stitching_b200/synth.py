"""Synthetic RealSense-like frames and the reference's calibration constants.

The reference's recordings (samples/*.bag) are Git-LFS pointers, so
tests and benchmarks run on seeded synthetic frames of the same shape: Z16 depth
in millimetres with 1/16 holes, RGB8 colour with stride 3*W.
"""
from __future__ import annotations

import numpy as np

# /root/reference/src/pcs-camera-optimized.cpp:64-67 (tf_mat, row-major 4x4)
TF_CAMERA = np.array([
    -0.99977970, 0.00926272, 0.01883480, 0.00000000,
    -0.01638983, 0.21604544, -0.97624574, 3.41600000,
    -0.01311186, -0.97633937, -0.21584603, 1.80200000,
    0.00000000, 0.00000000, 0.00000000, 1.00000000], dtype=np.float32)

# /root/reference/src/pcs-multicamera-optimized.cpp:417-455 (transform[0..7], row-major)
TF_STITCH = np.array([
    [-0.69888007, -0.32213748, 0.63858757, -2.22900000, -0.71520905, 0.32290986, -0.61984291,
     2.91800000, -0.00653159, -0.88991947, -0.45607091, 0.36400000, 0, 0, 0, 1],
    [-0.96127595, 0.09045863, -0.26031862, 0.31700000, 0.27558764, 0.31552831, -0.90801615,
     2.83300000, 0.00000000, -0.94459469, -0.32823906, 0.38100000, 0, 0, 0, 1],
    [-0.63305575, 0.28270490, -0.72063747, 2.80300000, 0.77409926, 0.22724638, -0.59087175,
     2.05500000, -0.00328008, -0.93189968, -0.36270128, 0.42100000, 0, 0, 0, 1],
    [0.17021299, 0.28598815, -0.94299433, 2.51000000, 0.98527137, -0.03349883, 0.16768470,
     -0.27300000, 0.01636663, -0.95764743, -0.28747787, 0.35900000, 0, 0, 0, 1],
    [0.72625904, 0.26139935, -0.63578155, 1.90900000, 0.68735231, -0.26305364, 0.67701520,
     -2.81700000, 0.00972668, -0.92869433, -0.37071853, 0.37900000, 0, 0, 0, 1],
    [0.98744750, 0.00686296, 0.15779838, -0.57400000, -0.14665062, -0.33120318, 0.93209337,
     -2.69700000, 0.05866025, -0.94353450, -0.32603930, 0.30900000, 0, 0, 0, 1],
    [0.67295609, 0.40193638, 0.62094867, -2.97300000, -0.35777412, -0.55787451, 0.74884826,
     -0.41700000, 0.64740079, -0.72610136, -0.23162261, 0.43400000, 0, 0, 0, 1],
    [0.08929624, -0.21535297, 0.97244500, -2.95700000, -0.67610010, -0.73004840, -0.09958907,
     -0.33900000, 0.73137872, -0.64857723, -0.21079074, 0.33800000, 0, 0, 0, 1],
], dtype=np.float32)

IDENTITY = np.eye(4, dtype=np.float32).reshape(-1)

# D435-like RGB baseline: depth -> colour translation of 15 mm along x
D2C_BASELINE = (0.015, 0.0, 0.0)




def rotation_colmajor(rx: float, ry: float, rz: float):
    """rs2_extrinsics-style rotation (column-major 3x3, float32) from small Euler angles in radians."""
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    rxm = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    rym = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rzm = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    r = (rzm @ rym @ rxm).astype(np.float32)
    return tuple(float(v) for v in r.T.reshape(-1))


# A factory-calibration-like depth -> colour rotation (a few tenths of a degree about every axis): any real
# D4xx has one; the reference's recordings are 1280x720 Z16 + 1920x1080 RGB8 (src/pcs-camera-grab-frames.cpp:69-70)
D2C_ROTATION_SMALL = rotation_colmajor(0.004, -0.003, 0.005)


def seed_for(cam: int, frame: int) -> int:
    return 0xC0FFEE ^ (cam << 16) ^ frame


#: depth noise of the default scene (mm): a 0.8 m thick cloud per camera, ~0.84 voxels per point at the 10 mm leaf --
#: the worst case for the voxel merge.  SMOOTH_NOISE_MM is a surface as a depth camera sees it (several points per voxel).
NOISE_MM = 400.0
SMOOTH_NOISE_MM = 5.0


def depth_frame(w: int, h: int, cam: int = 0, frame: int = 0, lo: int = 300, hi: int = 6000,
                hole_p: float = 1.0 / 16.0, noise_mm: float = NOISE_MM) -> np.ndarray:
    """uint16[h, w]: a tilted plane plus noise (sigma noise_mm) clipped to [lo, hi] mm, zero with prob hole_p."""
    rng = np.random.default_rng(seed_for(cam, frame))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = rng.uniform(1500, 3500)
    plane = base + rng.uniform(-1.5, 1.5) * (xx - w / 2) + rng.uniform(-1.5, 1.5) * (yy - h / 2)
    z = plane + rng.normal(0.0, noise_mm, size=(h, w)).astype(np.float32)
    z = np.clip(z, lo, hi).astype(np.uint16)
    z[rng.random((h, w)) < hole_p] = 0
    return z


def color_frame(w: int, h: int, cam: int = 0, frame: int = 0, stride: int | None = None,
                bpp: int = 3) -> np.ndarray:
    """uint8[h, stride]: uniform random bytes (RGB8, stride >= w*bpp)."""
    stride = w * bpp if stride is None else stride
    rng = np.random.default_rng(seed_for(cam, frame) ^ 0x5EED)
    return rng.integers(0, 256, size=(h, stride), dtype=np.uint8)


def frames(w: int, h: int, n_cams: int, n_frames: int, cw: int | None = None,
           ch: int | None = None):
    """depth uint16[n_cams, n_frames, h, w], colour uint8[n_cams, n_frames, ch, cw*3]."""
    cw = w if cw is None else cw
    ch = h if ch is None else ch
    d = np.stack([np.stack([depth_frame(w, h, c, f) for f in range(n_frames)])
                  for c in range(n_cams)])
    c = np.stack([np.stack([color_frame(cw, ch, c_, f) for f in range(n_frames)])
                  for c_ in range(n_cams)])
    return d, c
