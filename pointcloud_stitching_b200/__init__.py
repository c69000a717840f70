"""pointcloud_stitching_b200 -- B200-native hot path of conix-center/pointcloud_stitching.

The product is ``libpcs_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/pcs_b200.h``).  This module is the thin ctypes binding used by the
tests and by ``bench.py``; it holds no compute and has no CPU fallback: the
library is opened on first use of ``lib`` (so that ``synth`` / ``calibration`` can
be imported by processes that must not map the product, e.g. the CPU reference arm
of ``bench.py``); if it is missing that first use raises ImportError, and without a
CUDA device ``Context()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from . import synth  # noqa: F401  (synthetic frames + the reference's calibration constants)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# PCS_B200_LIB: an alternative build of the same library (A/B experiments); default is the in-tree build
LIB_PATH = os.environ.get("PCS_B200_LIB") or os.path.join(HERE, "libpcs_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "pcs_b200.h")

RECORD_BYTES = 10
HEADER_BYTES = 4
CAMERA_BUF_SHORTS = 5000000

PCS_OK, PCS_ERR_INVALID, PCS_ERR_CUDA, PCS_ERR_NOMEM, PCS_ERR_UNSUPPORTED, PCS_ERR_CAPACITY = (
    0, -1, -2, -3, -4, -5)


class PcsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("pcs_b200 error %d: %s" % (status, message))
        self.status = status


class Intrinsics(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ppx", C.c_float), ("ppy", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("model", C.c_int32), ("coeffs", C.c_float * 5)]


class StreamDesc(C.Structure):
    _fields_ = [("depth", Intrinsics), ("color", Intrinsics), ("d2c_rotation", C.c_float * 9),
                ("d2c_translation", C.c_float * 3), ("depth_scale", C.c_float),
                ("color_bpp", C.c_int32), ("color_stride", C.c_int32), ("tf", C.c_float * 16),
                ("cutoff", C.c_int32), ("z_lo", C.c_float), ("z_hi", C.c_float), ("x_lo", C.c_float),
                ("x_hi", C.c_float), ("cutoff_lane_reversed", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_streams", C.c_int32), ("kernel_variant", C.c_int32),
                ("voxel_variant", C.c_int32)]


class IpcHandle(C.Structure):
    _fields_ = [("reserved", C.c_uint8 * 64), ("offset", C.c_uint64), ("device", C.c_uint64)]


class ShardPeers(C.Structure):
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("inbox_dev", C.c_void_p * 8), ("cursor_dev", C.c_void_p * 8),
                ("zhist_dev", C.c_void_p * 8), ("capacity_records", C.c_int64)]


class FrameJob(C.Structure):
    _fields_ = [("stream", C.c_int32), ("reserved", C.c_int32), ("z16_dev", C.c_void_p),
                ("color_dev", C.c_void_p), ("payload_dev", C.c_void_p), ("xyzrgb_dev", C.c_void_p),
                ("count_dev", C.c_void_p)]


def stream_desc(dw, dh, cw=None, ch=None, tf=None, translation=(0.0, 0.0, 0.0), rotation=None,
                depth_scale=0.001, bpp=3, stride=None, cutoff=False, lane_reversed=True,
                dfx=None, dfy=None, dppx=None, dppy=None, cfx=None, cfy=None, cppx=None, cppy=None,
                depth_distortion=None, color_distortion=None):
    """A pcs_stream_desc with the SURVEY s8(d) synthetic calibration as defaults
    (f = W/2, principal point at the image centre, RGB8 colour of the depth size)."""
    cw = dw if cw is None else cw
    ch = dh if ch is None else ch
    d = StreamDesc()
    d.depth = Intrinsics(dw, dh, (dw - 1) / 2 if dppx is None else dppx,
                         (dh - 1) / 2 if dppy is None else dppy, dw / 2 if dfx is None else dfx,
                         dw / 2 if dfy is None else dfy)
    d.color = Intrinsics(cw, ch, (cw - 1) / 2 if cppx is None else cppx,
                         (ch - 1) / 2 if cppy is None else cppy, cw / 2 if cfx is None else cfx,
                         cw / 2 if cfy is None else cfy)
    if depth_distortion is not None:      # rs2 inverse Brown-Conrady: k1, k2, p1, p2, k3
        d.depth.model, d.depth.coeffs = 2, (C.c_float * 5)(*depth_distortion)
    if color_distortion is not None:      # rs2 modified Brown-Conrady
        d.color.model, d.color.coeffs = 1, (C.c_float * 5)(*color_distortion)
    d.d2c_rotation = (C.c_float * 9)(*((1, 0, 0, 0, 1, 0, 0, 0, 1) if rotation is None else rotation))
    d.d2c_translation = (C.c_float * 3)(*translation)
    d.depth_scale = depth_scale
    d.color_bpp = bpp
    d.color_stride = cw * bpp if stride is None else stride
    d.tf = (C.c_float * 16)(*(synth.IDENTITY if tf is None else np.asarray(tf, np.float32).reshape(-1)))
    d.cutoff = int(cutoff)
    # src/pcs-camera-optimized.cpp:398-401
    d.z_lo, d.z_hi, d.x_lo, d.x_hi = 0.0, 1.5, -2.0, 2.0
    d.cutoff_lane_reversed = int(lane_reversed)
    return d


def declared_symbols():
    """Every function include/pcs_b200.h declares."""
    with open(HEADER_PATH) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pcs_b200_\w+)\s*\(", text)))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    sig = {
        "pcs_b200_abi_version": (C.c_int, []),
        "pcs_b200_status_string": (C.c_char_p, [C.c_int]),
        "pcs_b200_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "pcs_b200_destroy": (None, [vp]),
        "pcs_b200_last_error": (C.c_char_p, [vp]),
        "pcs_b200_set_stream": (C.c_int, [vp, C.c_int, C.POINTER(StreamDesc)]),
        "pcs_b200_send_xyzrgb": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int]),
        "pcs_b200_send_xyzrgb_begin": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int]),
        "pcs_b200_send_xyzrgb_end": (C.c_int, [vp, C.c_int]),
        "pcs_b200_host_alloc": (vp, [vp, C.c_size_t]),
        "pcs_b200_host_free": (None, [vp, vp]),
        "pcs_b200_pack_from_vertices": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp]),
        "pcs_b200_pack_from_vertices_dev": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]),
        "pcs_b200_batch_create": (C.c_int, [vp, C.POINTER(FrameJob), C.c_int, C.POINTER(vp)]),
        "pcs_b200_batch_create_fanout": (C.c_int, [vp, C.POINTER(FrameJob), C.c_int, vp, C.c_size_t,
                                                   C.POINTER(vp), C.c_int, C.POINTER(vp)]),
        "pcs_b200_enable_peer": (C.c_int, [vp, C.c_int]),
        "pcs_b200_ipc_export": (C.c_int, [vp, vp, C.POINTER(IpcHandle)]),
        "pcs_b200_ipc_open": (C.c_int, [vp, C.POINTER(IpcHandle), C.POINTER(vp)]),
        "pcs_b200_ipc_close": (C.c_int, [vp, vp]),
        "pcs_b200_batch_run": (C.c_int, [vp, vp, vp]),
        "pcs_b200_batch_destroy": (None, [vp, vp]),
        "pcs_b200_batch_launches": (C.c_int, [vp]),
        "pcs_b200_stitch_raw_dev": (C.c_int, [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, vp, C.c_size_t, vp]),
        "pcs_b200_stitch_raw": (C.c_int, [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, vp, C.c_size_t]),
        "pcs_b200_stitch_pcl_dev": (C.c_int, [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, vp, vp, C.c_size_t, vp, vp]),
        "pcs_b200_stitch_pcl": (C.c_int, [vp, C.POINTER(vp), i32p, C.c_int, C.c_int, vp, vp, C.c_size_t]),
        "pcs_b200_stitch_frames_begin": (C.c_int, [vp, C.c_int, C.c_int, i32p, C.POINTER(vp), C.POINTER(vp), C.c_int,
                                                   vp, C.c_size_t]),
        "pcs_b200_stitch_frames_end": (C.c_int, [vp, C.c_int]),
        "pcs_b200_stitch_frames": (C.c_int, [vp, C.c_int, i32p, C.POINTER(vp), C.POINTER(vp), C.c_int, vp, C.c_size_t]),
        "pcs_b200_voxel_merge_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
        "pcs_b200_voxel_merge": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
        "pcs_b200_voxel_merge_async_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp]),
        "pcs_b200_voxel_merge_slab_async_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
        "pcs_b200_voxel_merge_counted_async_dev": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp]),
        "pcs_b200_shard_zbins": (C.c_int, [C.c_int]),
        "pcs_b200_shard_hist_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp]),
        "pcs_b200_shard_plan_dev": (C.c_int, [vp, C.POINTER(ShardPeers), C.c_int, vp, vp, vp]),
        "pcs_b200_shard_scatter_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, C.POINTER(ShardPeers), vp, vp]),
        "pcs_b200_voxel_slab_plan_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, i32p, i32p, vp]),
        "pcs_b200_voxel_merge_slab_dev": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
        "pcs_b200_cloud_to_ply_rows_dev": (C.c_int, [vp, vp, C.c_int, vp, vp]),
        "pcs_b200_save_ply": (C.c_int, [vp, vp, C.c_int, C.c_char_p]),
        "pcs_b200_synchronize": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


class _LazyLib:
    """dlopens libpcs_b200.so on first attribute access (never at import of the package)."""

    _real = None

    def __getattr__(self, name):
        real = object.__getattribute__(self, "_real")
        if real is None:
            real = _load()
            object.__setattr__(self, "_real", real)
        return getattr(real, name)

    @property
    def loaded(self):
        return object.__getattribute__(self, "_real") is not None


lib = _LazyLib()


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class Batch:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    @property
    def launches(self):
        return lib.pcs_b200_batch_launches(self.handle)

    def run(self, cuda_stream=0):
        self.ctx._check(lib.pcs_b200_batch_run(self.ctx.handle, self.handle, C.c_void_p(cuda_stream)))

    def close(self):
        if self.handle and lib is not None and self.ctx.handle:
            lib.pcs_b200_batch_destroy(self.ctx.handle, self.handle)
        self.handle = None

    __del__ = close


class Context:
    """pcs_ctx: one per process per GPU."""

    def __init__(self, device=0, max_streams=8, kernel_variant=0, voxel_variant=0):
        cfg = Config(device, max_streams, kernel_variant, voxel_variant)
        h = C.c_void_p()
        rc = lib.pcs_b200_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise PcsError(rc, (lib.pcs_b200_last_error(None) or b"").decode())
        self.handle = h
        self.descs = {}
        self._pinned = []

    def close(self):
        if getattr(self, "handle", None) and lib is not None:   # lib is None at interpreter shutdown
            for p in self._pinned:
                lib.pcs_b200_host_free(self.handle, p)
            self._pinned = []
            lib.pcs_b200_destroy(self.handle)
            self.handle = None

    __del__ = close

    def _check(self, rc):
        if rc < 0:
            raise PcsError(rc, (lib.pcs_b200_last_error(self.handle) or b"").decode())
        return rc

    def set_stream(self, stream, desc: StreamDesc):
        self._check(lib.pcs_b200_set_stream(self.handle, stream, C.byref(desc)))
        self.descs[stream] = desc

    def host_alloc(self, nbytes, dtype=np.uint8):
        """Pinned host memory as a numpy array (freed with the context)."""
        p = lib.pcs_b200_host_alloc(self.handle, nbytes)
        if not p:
            raise PcsError(PCS_ERR_NOMEM, (lib.pcs_b200_last_error(self.handle) or b"").decode())
        self._pinned.append(p)
        buf = (C.c_uint8 * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8).view(dtype)

    # ---- camera side, host buffers
    def new_camera_buffer(self, pinned=False):
        if pinned:
            b = self.host_alloc(CAMERA_BUF_SHORTS * 2, np.int16)
            b[:] = 0x5A5A
            return b
        return np.full(CAMERA_BUF_SHORTS, 0x5A5A, np.int16)

    def send_xyzrgb(self, stream, z16, color, buffer=None, write_header=False):
        """sendXYZRGBPointcloud: returns (payload_bytes, buffer int16[5 000 000])."""
        z16 = np.ascontiguousarray(z16, np.uint16)
        color = np.ascontiguousarray(color, np.uint8)
        if buffer is None:
            buffer = self.new_camera_buffer()
        size = self._check(lib.pcs_b200_send_xyzrgb(self.handle, stream, _np_ptr(z16), _np_ptr(color),
                                                    _np_ptr(buffer), int(write_header)))
        return size, buffer

    def send_begin(self, stream, z16, color, buffer, write_header=False):
        self._check(lib.pcs_b200_send_xyzrgb_begin(self.handle, stream, _np_ptr(z16), _np_ptr(color),
                                                   _np_ptr(buffer), int(write_header)))

    def send_end(self, stream):
        return self._check(lib.pcs_b200_send_xyzrgb_end(self.handle, stream))

    def pack_from_vertices(self, stream, xyz, uv, color):
        """copyPointCloudXYZRGBToBufferSIMD: returns records int16[count, 5]."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        uv = np.ascontiguousarray(uv, np.float32)
        color = np.ascontiguousarray(color, np.uint8)
        n = xyz.shape[0]
        out = np.zeros((n, 5), np.int16)
        cnt = self._check(lib.pcs_b200_pack_from_vertices(self.handle, stream, _np_ptr(xyz), _np_ptr(uv),
                                                          n, _np_ptr(color), _np_ptr(out)))
        return out[:cnt]

    # ---- camera side, batched / device resident
    @staticmethod
    def _job_array(jobs):
        arr = (FrameJob * len(jobs))()
        for i, j in enumerate(jobs):
            j = tuple(j) + (None,) * (7 - len(j))
            arr[i] = FrameJob(j[0], int(j[6] or 0), j[1], j[2], j[3], j[4], j[5])
        return arr

    def batch(self, jobs):
        """jobs: iterable of (stream, z16_ptr, color_ptr, payload_ptr[, xyzrgb_ptr[, count_ptr[, flags]]])."""
        h = C.c_void_p()
        self._check(lib.pcs_b200_batch_create(self.handle, self._job_array(jobs), len(jobs), C.byref(h)))
        return Batch(self, h)

    def batch_fanout(self, jobs, local_base, local_bytes, peer_bases):
        """Fused K1 + all-gather: records also go to the same offset of every peer mirror buffer."""
        h = C.c_void_p()
        peers = (C.c_void_p * max(1, len(peer_bases)))(*peer_bases)
        self._check(lib.pcs_b200_batch_create_fanout(self.handle, self._job_array(jobs), len(jobs), local_base,
                                                     local_bytes, peers, len(peer_bases), C.byref(h)))
        return Batch(self, h)

    def enable_peer(self, peer_device):
        """Single-process multi-GPU: let this context's kernels touch cudaMalloc memory of peer_device."""
        self._check(lib.pcs_b200_enable_peer(self.handle, peer_device))

    def ipc_export(self, dev_ptr):
        """bytes (80) that another process passes to ``ipc_open`` to map this device memory."""
        h = IpcHandle()
        self._check(lib.pcs_b200_ipc_export(self.handle, dev_ptr, C.byref(h)))
        return bytes(h)

    def ipc_open(self, handle_bytes):
        h = IpcHandle.from_buffer_copy(handle_bytes)
        p = C.c_void_p()
        self._check(lib.pcs_b200_ipc_open(self.handle, C.byref(h), C.byref(p)))
        return p.value

    def ipc_close(self, dev_ptr):
        self._check(lib.pcs_b200_ipc_close(self.handle, dev_ptr))

    def pack_from_vertices_dev(self, stream, xyz_ptr, uv_ptr, n, color_ptr, payload_ptr, count_ptr=None,
                               cuda_stream=0):
        return self._check(lib.pcs_b200_pack_from_vertices_dev(
            self.handle, stream, xyz_ptr, uv_ptr, n, color_ptr, payload_ptr, count_ptr, cuda_stream))

    # ---- stitch side
    @staticmethod
    def _cam_arrays(ptrs, n_shorts):
        n = len(ptrs)
        return (C.c_void_p * n)(*ptrs), (C.c_int32 * n)(*n_shorts)

    def stitch_raw_dev(self, ptrs, n_shorts, downsample, stitched_ptr, cap, cuda_stream=0):
        a, ns = self._cam_arrays(ptrs, n_shorts)
        return self._check(lib.pcs_b200_stitch_raw_dev(self.handle, a, ns, len(ptrs), downsample,
                                                       stitched_ptr, cap, cuda_stream))

    def stitch_pcl_dev(self, ptrs, n_shorts, downsample, transforms, stitched_ptr, cap, cloud32_ptr=None,
                       cuda_stream=0):
        a, ns = self._cam_arrays(ptrs, n_shorts)
        tf = np.ascontiguousarray(transforms, np.float32).reshape(-1)
        return self._check(lib.pcs_b200_stitch_pcl_dev(self.handle, a, ns, len(ptrs), downsample,
                                                       _np_ptr(tf), stitched_ptr, cap, cloud32_ptr,
                                                       cuda_stream))

    def _stitch_host(self, payloads, downsample, transforms):
        bufs = [np.ascontiguousarray(p, np.int16).reshape(-1) for p in payloads]
        a, ns = self._cam_arrays([b.ctypes.data for b in bufs], [b.size for b in bufs])
        cap = sum(b.size for b in bufs) * 2 + 4
        out = np.zeros(cap, np.uint8)
        if transforms is None:
            size = lib.pcs_b200_stitch_raw(self.handle, a, ns, len(bufs), downsample, _np_ptr(out), cap)
        else:
            tf = np.ascontiguousarray(transforms, np.float32).reshape(-1)
            size = lib.pcs_b200_stitch_pcl(self.handle, a, ns, len(bufs), downsample, _np_ptr(tf),
                                           _np_ptr(out), cap)
        self._check(size)
        return out[: size + 4]

    def stitch_raw(self, payloads, downsample=1):
        """sendStitchToUnity's concat: returns [int32 bytes][records] as uint8."""
        return self._stitch_host(payloads, downsample, None)

    def stitch_pcl(self, payloads, transforms, downsample=1):
        """unpack -> transform -> += -> repack: returns [int32 bytes][records] as uint8."""
        return self._stitch_host(payloads, downsample, transforms)

    # ---- camera + stitch side in one call (host frames in, the reference's stitched buffer out)
    STITCH_SLOTS = 4

    def stitch_frames_begin(self, slot, streams, z16_frames, color_frames, stitched, downsample=1):
        """Enqueue one stitched frame on pipeline `slot`: host z16 + colour frames of `streams` (stitched
        order) in, `[int32 bytes][records]` into the host array `stitched` (uint8) once ``_end`` returns."""
        n = len(streams)
        ids = (C.c_int32 * n)(*streams)
        zs = (C.c_void_p * n)(*[z.ctypes.data for z in z16_frames])
        cs = (C.c_void_p * n)(*[c.ctypes.data for c in color_frames])
        self._check(lib.pcs_b200_stitch_frames_begin(self.handle, slot, n, ids, zs, cs, downsample,
                                                     _np_ptr(stitched), stitched.nbytes))

    def stitch_frames_end(self, slot):
        return self._check(lib.pcs_b200_stitch_frames_end(self.handle, slot))

    def stitch_frames(self, streams, z16_frames, color_frames, downsample=1, stitched=None):
        """sendXYZRGBPointcloud on every camera + sendStitchToUnity's concat: returns uint8 [int32][records]."""
        z16_frames = [np.ascontiguousarray(z, np.uint16) for z in z16_frames]
        color_frames = [np.ascontiguousarray(c, np.uint8) for c in color_frames]
        if stitched is None:
            stitched = np.zeros(4 + 10 * sum(z.size for z in z16_frames), np.uint8)
        self.stitch_frames_begin(0, streams, z16_frames, color_frames, stitched, downsample)
        size = self.stitch_frames_end(0)
        return stitched[: size + 4]

    def voxel_merge(self, records, leaf_mm=10):
        rec = np.ascontiguousarray(records, np.int16).reshape(-1, 5)
        out = np.zeros_like(rec)
        n = self._check(lib.pcs_b200_voxel_merge(self.handle, _np_ptr(rec), rec.shape[0], leaf_mm,
                                                 _np_ptr(out)))
        return out[:n]

    def voxel_merge_dev(self, records_ptr, n, leaf_mm, out_ptr, cuda_stream=0):
        return self._check(lib.pcs_b200_voxel_merge_dev(self.handle, records_ptr, n, leaf_mm, out_ptr,
                                                        cuda_stream))

    def voxel_merge_async_dev(self, records_ptr, n, leaf_mm, out_ptr, count_ptr, cuda_stream=0):
        """Enqueue only: the voxel count (or a negative status) lands in the device int32 at count_ptr."""
        self._check(lib.pcs_b200_voxel_merge_async_dev(self.handle, records_ptr, n, leaf_mm, out_ptr, count_ptr,
                                                       cuda_stream))

    def voxel_merge_slab_async_dev(self, records_ptr, n, leaf_mm, kz_lo, kz_hi, out_ptr, count_ptr, cuda_stream=0):
        self._check(lib.pcs_b200_voxel_merge_slab_async_dev(self.handle, records_ptr, n, leaf_mm, kz_lo, kz_hi, out_ptr,
                                                            count_ptr, cuda_stream))

    def voxel_merge_counted_async_dev(self, records_ptr, n_max, n_ptr, leaf_mm, out_ptr, count_ptr, cuda_stream=0):
        """Merge of an inbox whose fill count is the device int32 at n_ptr (<= n_max)."""
        self._check(lib.pcs_b200_voxel_merge_counted_async_dev(self.handle, records_ptr, n_max, n_ptr, leaf_mm, out_ptr,
                                                               count_ptr, cuda_stream))

    # ---- multi-GPU: shard by voxel-key range before the exchange
    def shard_hist_dev(self, records_ptr, n, leaf_mm, zhist_ptr, cursor_ptr, cuda_stream=0):
        self._check(lib.pcs_b200_shard_hist_dev(self.handle, records_ptr, n, leaf_mm, zhist_ptr, cursor_ptr, cuda_stream))

    def shard_plan_dev(self, peers: "ShardPeers", leaf_mm, splits_ptr, zslab_ptr, cuda_stream=0):
        self._check(lib.pcs_b200_shard_plan_dev(self.handle, C.byref(peers), leaf_mm, splits_ptr, zslab_ptr, cuda_stream))

    def shard_scatter_dev(self, records_ptr, n, leaf_mm, zslab_ptr, peers: "ShardPeers", err_ptr, cuda_stream=0):
        self._check(lib.pcs_b200_shard_scatter_dev(self.handle, records_ptr, n, leaf_mm, zslab_ptr, C.byref(peers), err_ptr,
                                                   cuda_stream))

    def voxel_slab_plan_dev(self, records_ptr, n, leaf_mm, n_slabs, cuda_stream=0):
        """(kz_splits[n_slabs + 1], slab_points[n_slabs]): equal-population cuts of the grid along z."""
        splits, pts = (C.c_int32 * (n_slabs + 1))(), (C.c_int32 * n_slabs)()
        self._check(lib.pcs_b200_voxel_slab_plan_dev(self.handle, records_ptr, n, leaf_mm, n_slabs, splits, pts,
                                                     cuda_stream))
        return list(splits), list(pts)

    def voxel_merge_slab_dev(self, records_ptr, n, leaf_mm, kz_lo, kz_hi, out_ptr, cuda_stream=0):
        return self._check(lib.pcs_b200_voxel_merge_slab_dev(self.handle, records_ptr, n, leaf_mm, kz_lo, kz_hi,
                                                             out_ptr, cuda_stream))

    def cloud_to_ply_rows_dev(self, cloud32_ptr, n, rows_ptr, cuda_stream=0):
        return self._check(lib.pcs_b200_cloud_to_ply_rows_dev(self.handle, cloud32_ptr, n, rows_ptr, cuda_stream))

    def save_ply(self, cloud32_ptr, n, path):
        """pcl::io::savePLYFileBinary of the stitched cloud (32-byte PCL points on the device)."""
        return self._check(lib.pcs_b200_save_ply(self.handle, cloud32_ptr, n, os.fsencode(path)))

    def synchronize(self, cuda_stream=0):
        self._check(lib.pcs_b200_synchronize(self.handle, cuda_stream))
