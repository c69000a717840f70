"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle.

May be imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and by nothing
else.  The product package ``pointcloud_stitching_b200`` never imports it and
has no CPU fallback.

Two libraries:

* ``restatement`` -- ``oracle/pcs_oracle.c`` (plain C, this repo), built into
  ``oracle/_build/libpcs_oracle.so``.  Always available (gcc is in the image).
* ``ref_camera`` / ``ref_client`` / ``ref_optimized`` -- the reference's own
  translation units compiled unmodified from ``/root/reference`` against stub
  headers (``oracle/Makefile`` target ``ref``) into ``oracle/_ref/``.  Built in
  the dev container; the prebuilt files travel to the GPU box.  ``None`` when
  absent.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF = os.path.join(HERE, "_ref")


class Intrinsics(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("ppx", C.c_float), ("ppy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("model", C.c_int), ("coeffs", C.c_float * 5)]


class Calib(C.Structure):
    _fields_ = [("depth", Intrinsics), ("color", Intrinsics), ("rotation", C.c_float * 9),
                ("translation", C.c_float * 3), ("depth_scale", C.c_float)]


PCLPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4"), ("b", "u1"),
                     ("g", "u1"), ("r", "u1"), ("a", "u1"), ("pad", "<u4", (3,))])
assert PCLPOINT.itemsize == 32


def make_calib(dw, dh, cw=None, ch=None, dfx=None, dfy=None, dppx=None, dppy=None, cfx=None,
               cfy=None, cppx=None, cppy=None, rotation=None, translation=(0.0, 0.0, 0.0),
               depth_scale=0.001, depth_distortion=None, color_distortion=None):
    """Calibration with the SURVEY s8(d) defaults: f = W/2, pp = ((W-1)/2, (H-1)/2)."""
    cw = dw if cw is None else cw
    ch = dh if ch is None else ch
    c = Calib()
    c.depth = Intrinsics(dw, dh, (dw - 1) / 2 if dppx is None else dppx,
                         (dh - 1) / 2 if dppy is None else dppy, dw / 2 if dfx is None else dfx,
                         dw / 2 if dfy is None else dfy)
    c.color = Intrinsics(cw, ch, (cw - 1) / 2 if cppx is None else cppx,
                         (ch - 1) / 2 if cppy is None else cppy, cw / 2 if cfx is None else cfx,
                         cw / 2 if cfy is None else cfy)
    if depth_distortion is not None:          # inverse Brown-Conrady (k1, k2, p1, p2, k3), applied on deprojection
        c.depth.model, c.depth.coeffs = 2, (C.c_float * 5)(*depth_distortion)
    if color_distortion is not None:          # modified Brown-Conrady, applied on projection
        c.color.model, c.color.coeffs = 1, (C.c_float * 5)(*color_distortion)
    rot = (1, 0, 0, 0, 1, 0, 0, 0, 1) if rotation is None else tuple(rotation)
    c.rotation = (C.c_float * 9)(*rot)
    c.translation = (C.c_float * 3)(*translation)
    c.depth_scale = depth_scale
    return c


def build(ref: bool = True) -> None:
    """Compile the restatement and (when /root/reference exists) the reference TUs."""
    targets = ["all"] + (["ref"] if ref else [])
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True, env=env)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Restatement:
    """oracle/pcs_oracle.c"""

    def __init__(self):
        path = os.path.join(BUILD, "libpcs_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        L = self.lib
        fp, u8p, u16p, i16p = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16),
                               C.POINTER(C.c_int16))
        L.pcs_oracle_deproject.argtypes = [C.POINTER(Calib), u16p, fp, fp, C.c_int]
        L.pcs_oracle_deproject.restype = None
        L.pcs_oracle_pack_simd.argtypes = [fp, fp, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           fp, C.c_int, i16p]
        L.pcs_oracle_transform_points.argtypes = [fp, C.c_int, fp, fp]
        L.pcs_oracle_transform_points.restype = None
        L.pcs_oracle_send.argtypes = [fp, fp, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, fp,
                                      C.c_int, C.c_int, i16p]
        L.pcs_oracle_concat.argtypes = [C.POINTER(i16p), C.POINTER(C.c_int), C.c_int, C.c_int, i16p]
        L.pcs_oracle_unpack.argtypes = [i16p, C.c_int, C.c_int, C.c_void_p]
        L.pcs_oracle_transform_cloud.argtypes = [C.c_void_p, C.c_int, fp]
        L.pcs_oracle_transform_cloud.restype = None
        L.pcs_oracle_repack.argtypes = [C.c_void_p, C.c_int, i16p]
        L.pcs_oracle_voxel_merge.argtypes = [i16p, C.c_int, C.c_int, i16p]

    # -- camera side ---------------------------------------------------------
    def deproject(self, calib: Calib, z16: np.ndarray, threads: int = 1):
        z16 = np.ascontiguousarray(z16, dtype=np.uint16)
        n = calib.depth.width * calib.depth.height
        assert z16.size == n
        xyz = np.empty((n, 3), np.float32)
        uv = np.empty((n, 2), np.float32)
        self.lib.pcs_oracle_deproject(C.byref(calib), _ptr(z16, C.c_uint16), _ptr(xyz, C.c_float),
                                      _ptr(uv, C.c_float), threads)
        return xyz, uv

    def pack(self, xyz, uv, color, cw, ch, bpp, stride, tf, cutoff=False):
        xyz, uv, tf = _f32(xyz), _f32(uv), _f32(tf)
        color = np.ascontiguousarray(color, dtype=np.uint8)
        n = xyz.shape[0]
        out = np.zeros((n, 5), np.int16)
        cnt = self.lib.pcs_oracle_pack_simd(_ptr(xyz, C.c_float), _ptr(uv, C.c_float), n,
                                            _ptr(color, C.c_uint8), cw, ch, bpp, stride,
                                            _ptr(tf, C.c_float), int(cutoff), _ptr(out, C.c_int16))
        if cnt < 0:
            raise ValueError("pcs_oracle_pack_simd: n must be a multiple of 4")
        return out[:cnt]

    def transform_points(self, xyz, tf):
        xyz, tf = _f32(xyz), _f32(tf)
        out = np.empty_like(xyz)
        self.lib.pcs_oracle_transform_points(_ptr(xyz, C.c_float), xyz.shape[0], _ptr(tf, C.c_float),
                                             _ptr(out, C.c_float))
        return out

    def send(self, xyz, uv, color, cw, ch, bpp, stride, tf, cutoff=False, write_header=False,
             buffer=None):
        """Returns (payload_bytes, buffer as int16[5 000 000])."""
        xyz, uv, tf = _f32(xyz), _f32(uv), _f32(tf)
        color = np.ascontiguousarray(color, dtype=np.uint8)
        if buffer is None:
            buffer = np.full(5000000, 0x5A5A, np.int16)
        size = self.lib.pcs_oracle_send(_ptr(xyz, C.c_float), _ptr(uv, C.c_float), xyz.shape[0],
                                        _ptr(color, C.c_uint8), cw, ch, bpp, stride,
                                        _ptr(tf, C.c_float), int(cutoff), int(write_header),
                                        _ptr(buffer, C.c_int16))
        return size, buffer

    def frame(self, calib: Calib, z16, color, bpp, stride, tf, cutoff=False):
        """deproject o pack: the oracle of the fused kernel.  Returns records[n,5] int16."""
        xyz, uv = self.deproject(calib, z16)
        return self.pack(xyz, uv, color, calib.color.width, calib.color.height, bpp, stride, tf,
                         cutoff)

    # -- stitch side ---------------------------------------------------------
    def concat(self, payloads, downsample=1):
        """payloads: list of int16 arrays (records flattened).  Returns the stitched bytes
        [i32 size][records] as uint8."""
        bufs = [np.ascontiguousarray(p, dtype=np.int16).reshape(-1) for p in payloads]
        total = sum(-(-b.size // (5 * downsample)) * 5 for b in bufs)
        out = np.zeros(total + 2, np.int16)
        arr = (C.POINTER(C.c_int16) * len(bufs))(*[_ptr(b, C.c_int16) for b in bufs])
        ns = (C.c_int * len(bufs))(*[b.size for b in bufs])
        size = self.lib.pcs_oracle_concat(arr, ns, len(bufs), downsample, _ptr(out, C.c_int16))
        assert size == total * 2
        return out.view(np.uint8)[: size + 4]

    def unpack(self, records, downsample=1):
        rec = np.ascontiguousarray(records, dtype=np.int16).reshape(-1, 5)
        out = np.zeros(rec.shape[0] // downsample, PCLPOINT)
        n = self.lib.pcs_oracle_unpack(_ptr(rec, C.c_int16), rec.shape[0], downsample,
                                       out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def transform_cloud(self, pts, m):
        pts = np.array(pts, dtype=PCLPOINT, copy=True)
        m = _f32(m)
        self.lib.pcs_oracle_transform_cloud(pts.ctypes.data_as(C.c_void_p), pts.shape[0],
                                            _ptr(m, C.c_float))
        return pts

    def repack(self, pts):
        pts = np.ascontiguousarray(pts, dtype=PCLPOINT)
        out = np.zeros((pts.shape[0], 5), np.int16)
        n = self.lib.pcs_oracle_repack(pts.ctypes.data_as(C.c_void_p), pts.shape[0],
                                       _ptr(out, C.c_int16))
        return out[:n]

    def pcl_stitch(self, payloads, transforms, downsample=1):
        """unpack -> transform -> += -> repack for every camera in order
        (src/pcs-multicamera-optimized.cpp:354-382).  Returns [i32 size][records] bytes."""
        clouds = [self.transform_cloud(self.unpack(p, downsample), t)
                  for p, t in zip(payloads, transforms)]
        rec = self.repack(np.concatenate(clouds)) if clouds else np.zeros((0, 5), np.int16)
        out = np.zeros(rec.size * 2 + 4, np.uint8)
        out[:4] = np.frombuffer(np.int32(rec.size * 2).tobytes(), np.uint8)
        out[4:] = rec.reshape(-1).view(np.uint8)
        return out

    def voxel_merge(self, records, leaf_mm=10):
        rec = np.ascontiguousarray(records, dtype=np.int16).reshape(-1, 5)
        out = np.zeros_like(rec)
        n = self.lib.pcs_oracle_voxel_merge(_ptr(rec, C.c_int16), rec.shape[0], leaf_mm,
                                            _ptr(out, C.c_int16))
        if n < 0:
            raise ValueError("pcs_oracle_voxel_merge failed")
        return out[:n]


class RefCamera:
    """The reference's src/pcs-camera-optimized.cpp, compiled unmodified."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        L = self.lib
        fp, u8p, u16p, i16p = (C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16),
                               C.POINTER(C.c_int16))
        L.ref_pack.argtypes = [C.c_int, fp, fp, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, fp,
                               C.c_int, C.c_int, i16p]
        L.ref_send.argtypes = [fp, fp, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int,
                               C.c_int, i16p, C.c_int, u8p, C.c_int, C.POINTER(C.c_int)]
        L.ref_time_send.argtypes = [fp, fp, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, fp,
                                    C.c_int, C.c_int, i16p, C.c_int, C.POINTER(C.c_double)]
        L.ref_replay.argtypes = [u16p, u8p, C.c_int, C.POINTER(Calib), C.c_int, C.c_int, fp, C.c_int,
                                 C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p,
                                 C.c_int]

    def pack(self, xyz, uv, color, cw, ch, bpp, stride, tf, cutoff=False, simd=True, threads=1):
        xyz, uv, tf = _f32(xyz), _f32(uv), _f32(tf)
        color = np.ascontiguousarray(color, dtype=np.uint8)
        n = xyz.shape[0]
        assert n % 4 == 0
        out = np.zeros((n, 5), np.int16)
        cnt = self.lib.ref_pack(int(simd), _ptr(xyz, C.c_float), _ptr(uv, C.c_float), n,
                                _ptr(color, C.c_uint8), cw, ch, bpp, stride, _ptr(tf, C.c_float),
                                int(cutoff), threads, _ptr(out, C.c_int16))
        return out[:cnt]

    def send(self, xyz, uv, color, cw, ch, bpp, stride, tf, cutoff=False, threads=1, wire=False,
             buffer=None):
        """Returns (payload_bytes, buffer int16[5 000 000], wire bytes or None)."""
        xyz, uv, tf = _f32(xyz), _f32(uv), _f32(tf)
        color = np.ascontiguousarray(color, dtype=np.uint8)
        if buffer is None:
            buffer = np.full(5000000, 0x5A5A, np.int16)
        cap = 10 * xyz.shape[0] + 4
        wire_out = np.zeros(cap, np.uint8)
        wl = C.c_int(0)
        size = self.lib.ref_send(_ptr(xyz, C.c_float), _ptr(uv, C.c_float), xyz.shape[0],
                                 _ptr(color, C.c_uint8), cw, ch, bpp, stride, _ptr(tf, C.c_float),
                                 int(cutoff), threads, _ptr(buffer, C.c_int16), int(wire),
                                 _ptr(wire_out, C.c_uint8), cap, C.byref(wl))
        return size, buffer, (wire_out[: wl.value] if wire else None)

    def time_send(self, xyz, uv, color, cw, ch, bpp, stride, tf, iters, simd=True, threads=1):
        """Per-call milliseconds of sendXYZRGBPointcloud, timed like the reference does."""
        xyz, uv, tf = _f32(xyz), _f32(uv), _f32(tf)
        color = np.ascontiguousarray(color, dtype=np.uint8)
        buffer = np.zeros(5000000, np.int16)
        ms = np.zeros(iters, np.float64)
        self.lib.ref_time_send(_ptr(xyz, C.c_float), _ptr(uv, C.c_float), xyz.shape[0],
                               _ptr(color, C.c_uint8), cw, ch, bpp, stride, _ptr(tf, C.c_float),
                               int(simd), threads, _ptr(buffer, C.c_int16), iters,
                               _ptr(ms, C.c_double))
        return ms

    def replay(self, calib, depth_frames, color_frames, bpp, stride, tf, simd=True, threads=1):
        """Runs the reference main() replay loop; returns (avg_ms, calc_ms, log tail)."""
        d = np.ascontiguousarray(depth_frames, dtype=np.uint16)
        c = np.ascontiguousarray(color_frames, dtype=np.uint8)
        tf = _f32(tf)
        avg, calc = C.c_double(0), C.c_double(0)
        log = C.create_string_buffer(4096)
        self.lib.ref_replay(_ptr(d, C.c_uint16), _ptr(c, C.c_uint8), d.shape[0], C.byref(calib), bpp,
                            stride, _ptr(tf, C.c_float), int(simd), threads, C.byref(avg),
                            C.byref(calc), log, 4096)
        return avg.value, calc.value, log.value.decode(errors="replace")


class RefStitch:
    """src/pcs-multicamera-client.cpp (client=True) or src/pcs-multicamera-optimized.cpp."""

    def __init__(self, path, client):
        self.lib = C.CDLL(path)
        self.client = client
        L = self.lib
        fp, u8p, i16p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int16)
        L.ref_unpack.argtypes = [i16p, C.c_int, C.c_int, C.c_void_p]
        L.ref_repack.argtypes = [C.c_void_p, C.c_int, i16p]
        L.ref_pcl_stitch_1cam.argtypes = [i16p, C.c_int, C.c_int, fp, u8p, C.c_int]
        if client:
            L.ref_raw_stitch_1cam.argtypes = [i16p, C.c_int, C.c_int, u8p, C.c_int]

    def unpack(self, records, downsample=1):
        rec = np.ascontiguousarray(records, dtype=np.int16).reshape(-1, 5)
        assert rec.shape[0] % downsample == 0
        out = np.zeros(rec.shape[0] // downsample, PCLPOINT)
        n = self.lib.ref_unpack(_ptr(rec, C.c_int16), rec.shape[0], downsample,
                                out.ctypes.data_as(C.c_void_p))
        return out[:n]

    def repack(self, pts):
        pts = np.ascontiguousarray(pts, dtype=PCLPOINT)
        out = np.zeros((pts.shape[0], 5), np.int16)
        n = self.lib.ref_repack(pts.ctypes.data_as(C.c_void_p), pts.shape[0], _ptr(out, C.c_int16))
        return out[:n]

    def pcl_stitch_1cam(self, records, tf, downsample=1):
        rec = np.ascontiguousarray(records, dtype=np.int16).reshape(-1)
        tf = _f32(tf)
        out = np.zeros(rec.size * 2 + 4, np.uint8)
        n = self.lib.ref_pcl_stitch_1cam(_ptr(rec, C.c_int16), rec.size, downsample,
                                         _ptr(tf, C.c_float), _ptr(out, C.c_uint8), out.size)
        assert n >= 0
        return out[:n]

    def raw_stitch_1cam(self, records, downsample=1):
        assert self.client
        rec = np.ascontiguousarray(records, dtype=np.int16).reshape(-1)
        out = np.zeros(rec.size * 2 + 4, np.uint8)
        n = self.lib.ref_raw_stitch_1cam(_ptr(rec, C.c_int16), rec.size, downsample,
                                         _ptr(out, C.c_uint8), out.size)
        assert n >= 0
        return out[:n]


_cache = {}


def restatement() -> Restatement:
    if "r" not in _cache:
        _cache["r"] = Restatement()
    return _cache["r"]


def _ref(name, ctor):
    if name not in _cache:
        path = os.path.join(REF, "libpcs_ref_%s.so" % name)
        _cache[name] = ctor(path) if os.path.exists(path) else None
    return _cache[name]


def ref_camera():
    return _ref("camera", RefCamera)


def ref_client():
    return _ref("client", lambda p: RefStitch(p, True))


def ref_optimized():
    return _ref("optimized", lambda p: RefStitch(p, False))
