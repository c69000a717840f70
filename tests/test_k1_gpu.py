"""-m gpu parity tests, camera side: the CUDA path through the C ABI against
(a) the golden fixtures produced by the reference's own compiled functions and
(b) the oracle restatement on full-size seeded frames.  Bit-exact on all 10
bytes of every record; float XYZ within 1e-5 relative (north_star)."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import oracle  # noqa: E402
import pointcloud_stitching_b200 as pcs  # noqa: E402
from gpu_util import calib_and_desc, dev, run_batch, small_rotation  # noqa: E402
from pointcloud_stitching_b200 import synth  # noqa: E402

VARIANTS = [1, 2]  # 1 = direct, 2 = bulk-async pipelined


@pytest.fixture(scope="module")
def R():
    return oracle.restatement()


@pytest.fixture(scope="module", params=VARIANTS, ids=["direct", "pipelined"])
def ctx(request):
    c = pcs.Context(device=0, max_streams=8, kernel_variant=request.param)
    c.variant = request.param
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx0():
    c = pcs.Context(device=0, max_streams=8)
    yield c
    c.close()


# ------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", ["pack_96x64_identity", "pack_96x64_baseline", "pack_96x64_cutoff"])
def test_fused_kernel_vs_reference_golden(ctx0, name):
    g = load_golden(name)
    w, h = int(g["w"]), int(g["h"])
    cut = bool(g["cutoff"])
    ctx0.set_stream(0, pcs.stream_desc(w, h, tf=g["tf"], translation=tuple(g["translation"]), cutoff=cut))
    size, buf = ctx0.send_xyzrgb(0, g["z16"], g["color"])
    want = g["out_records"]
    assert size == want.shape[0] * 10
    assert np.array_equal(buf[2:2 + want.size].reshape(-1, 5), want)


@pytest.mark.parametrize("name", ["pack_96x64_identity", "pack_96x64_baseline", "pack_96x64_cutoff",
                                  "pack_adversarial", "pack_adversarial_cutoff"])
def test_from_vertices_vs_reference_golden(ctx0, name):
    g = load_golden(name)
    w, h = int(g["w"]), int(g["h"])
    stride = int(g["stride"]) if "stride" in g else w * 3
    ctx0.set_stream(1, pcs.stream_desc(w, h, tf=g["tf"], stride=stride, cutoff=bool(g["cutoff"])))
    got = ctx0.pack_from_vertices(1, g["xyz"], g["uv"], g["color"])
    assert np.array_equal(got, g["out_records"])


def test_camera_buffer_image_vs_reference_golden(ctx0):
    g = load_golden("send_64x32")
    w, h = int(g["w"]), int(g["h"])
    n5 = w * h * 5 + 64
    ctx0.set_stream(0, pcs.stream_desc(w, h, tf=g["tf"]))
    size, buf = ctx0.send_xyzrgb(0, g["z16"], g["color"], write_header=False)
    assert size == int(g["size"])
    assert np.array_equal(buf[:n5], g["out_head_nosend"])
    assert np.array_equal(buf[2499990:2500010], g["out_memset_edge"])
    assert np.all(buf[2500000:] == 0x5A5A)                      # memset stops at byte 5 000 000
    size, buf = ctx0.send_xyzrgb(0, g["z16"], g["color"], write_header=True)
    assert np.array_equal(buf[:n5], g["out_head_send"])
    assert np.array_equal(buf.view(np.uint8)[: size + 4], g["wire_bytes"])


# ------------------------------------------------------------------ full frames vs oracle
CASES = {
    "720p_aligned": dict(w=1280, h=720),
    "720p_baseline": dict(w=1280, h=720, translation=synth.D2C_BASELINE),
    "720p_rotated": dict(w=1280, h=720, translation=(0.015, -0.002, 0.001), rotation=small_rotation()),
    # a factory-calibration-sized rotation (~0.1 degree) and sub-millimetre y/z offsets: the windowed path
    "720p_rot_small": dict(w=1280, h=720, translation=(0.0148, 0.0002, 0.0003),
                           rotation=small_rotation(0.0015, -0.002, 0.001)),
    "480p_rot_small": dict(w=848, h=480, translation=(0.0148, -0.0003, 0.0002),
                           rotation=small_rotation(-0.002, 0.001, 0.0015)),
    "720p_translate_yz": dict(w=1280, h=720, translation=(0.015, 0.0004, -0.0006)),
    "480p_aligned": dict(w=848, h=480),
    "480p_baseline": dict(w=848, h=480, translation=synth.D2C_BASELINE),
    "720p_color1080p": dict(w=1280, h=720, cw=1920, ch=1080, translation=synth.D2C_BASELINE),
    # lens distortion (rs2_intrinsics.model / coeffs): modified Brown-Conrady on the colour sensor, inverse Brown-Conrady
    # on the depth sensor, both at once with a rotated extrinsic and colour of another size
    "480p_color_distortion": dict(w=848, h=480, translation=synth.D2C_BASELINE,
                                  color_distortion=(0.12, -0.25, 0.0012, -0.0008, 0.09)),
    "480p_depth_distortion": dict(w=848, h=480, depth_distortion=(-0.05, 0.07, 0.0005, 0.0011, -0.02)),
    "720p_both_distortions_rot_1080p": dict(w=1280, h=720, cw=1920, ch=1080, translation=(0.015, 0.001, -0.002),
                                            rotation=small_rotation(0.004, -0.003, 0.005),
                                            depth_distortion=(0.02, -0.03, 0.0004, -0.0006, 0.01),
                                            color_distortion=(0.1, -0.21, 0.001, 0.0015, 0.07)),
    "720p_rgba_padded": dict(w=1280, h=720, bpp=4, stride=1280 * 4 + 64, translation=synth.D2C_BASELINE),
    "odd_intrinsics": dict(w=640, h=360, translation=(0.02, 0.01, -0.005), dfx=381.7, dfy=380.9,
                           dppx=322.3, dppy=178.8, cfx=610.2, cfy=611.9, cppx=318.4, cppy=182.1),
    "tiny_8x1": dict(w=8, h=1),
    "narrow_24x5": dict(w=24, h=5, translation=synth.D2C_BASELINE),
    "small_64x4": dict(w=64, h=4),
    "small_128x6_baseline": dict(w=128, h=6, translation=synth.D2C_BASELINE),
    "wide_2048x16_baseline": dict(w=2048, h=16, translation=(0.05, 0, 0)),
}
# what the bulk-async pipelined kernel must accept (pipe_supports, pcs_k1_pipe.cuh); what it cannot take
# kernel_variant=2 must refuse loudly rather than fall back (more rotated rigs may qualify: the host sizes their
# windows from the calibration)
PIPELINED = {"720p_aligned", "720p_baseline", "480p_aligned", "480p_baseline", "small_64x4",
             "small_128x6_baseline", "wide_2048x16_baseline", "720p_rot_small", "480p_rot_small",
             "720p_translate_yz", "720p_color1080p"}


@pytest.mark.parametrize("case", list(CASES))
def test_fused_kernel_vs_oracle(ctx, R, case):
    kw = dict(CASES[case])
    w, h = kw.pop("w"), kw.pop("h")
    cw, ch = kw.pop("cw", w), kw.pop("ch", h)
    bpp, stride = kw.pop("bpp", 3), kw.pop("stride", None)
    stride = cw * bpp if stride is None else stride
    cal, desc = calib_and_desc(w, h, cw, ch, tf=synth.TF_STITCH[1], bpp=bpp, stride=stride, **kw)
    ctx.set_stream(0, desc)
    jobs = []
    for f in range(2):
        z = synth.depth_frame(w, h, 3, f)
        col = synth.color_frame(cw, ch, 3, f, stride=stride, bpp=bpp)
        jobs.append((0, z, col))
    try:
        got = run_batch(ctx, jobs, None)
    except pcs.PcsError as e:
        # kernel_variant = 2 refuses loudly what the pipelined kernel cannot take (never a silent fallback); the cases
        # listed in PIPELINED must run
        assert ctx.variant == 2 and case not in PIPELINED and e.status == pcs.PCS_ERR_UNSUPPORTED, (case, str(e))
        return
    for (_, z, col), (rec, _, _) in zip(jobs, got):
        want = R.frame(cal, z, col, bpp, stride, synth.TF_STITCH[1])
        assert rec.shape == want.shape
        bad = np.nonzero((rec != want).any(axis=1))[0]
        assert bad.size == 0, "first mismatches at points %s: got %s want %s" % (bad[:5], rec[bad[:5]], want[bad[:5]])


def test_windowed_taps_fall_back_to_global_loads(ctx, R):
    """Very near depths under a rotated calibration: most taps land outside the staged colour-row
    window (and many are clamped at the image border); the result must not change."""
    w, h = 1280, 720
    kw = dict(translation=(0.0148, 0.0002, 0.0003), rotation=small_rotation(0.0015, -0.002, 0.001))
    cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[3], **kw)
    ctx.set_stream(1, desc)
    z = synth.depth_frame(w, h, 8, 8, lo=1, hi=120)
    col = synth.color_frame(w, h, 8, 8)
    (rec, _, _), = run_batch(ctx, [(1, z, col)], None)
    assert np.array_equal(rec, R.frame(cal, z, col, 3, w * 3, synth.TF_STITCH[3]))


def test_extreme_depths(ctx, R):
    # z16 = 1 and 65535 everywhere, checkerboard with holes
    w, h = 256, 64
    for trans in [(0.0, 0.0, 0.0), synth.D2C_BASELINE]:
        cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[7], translation=trans)
        ctx.set_stream(2, desc)
        z = np.zeros((h, w), np.uint16)
        z[::2, ::2] = 65535
        z[1::2, 1::2] = 1
        z[::3, 1::4] = 32768
        col = synth.color_frame(w, h, 9, 9)
        (rec, _, _), = run_batch(ctx, [(2, z, col)], None)
        assert np.array_equal(rec, R.frame(cal, z, col, 3, w * 3, synth.TF_STITCH[7]))


SWEEPS = {
    "baseline": dict(translation=synth.D2C_BASELINE),
    # guarded taps (pcs_guard.h): every depth value meets the guard at every column, with the near depths the guard
    # hands to the exact chain, the taps that leave the staged window and the ones clamped at the frame's border
    "rotated": dict(translation=synth.D2C_BASELINE, rotation=synth.D2C_ROTATION_SMALL),
    "rotated_1080p": dict(cw=1920, ch=1080, translation=(0.0149, 0.0003, -0.0004), rotation=small_rotation(-0.003, 0.004, -0.006)),
}


@pytest.mark.parametrize("w,h,sweep", [(1280, 720, "baseline"), (848, 480, "baseline"), (1280, 720, "rotated"),
                                       (848, 480, "rotated"), (1280, 720, "rotated_1080p")])
def test_every_depth_value_at_every_column(ctx, R, w, h, sweep):
    """Exhaustive over the kernel's data-dependent arithmetic: every z16 in 0..65535 at every
    column (rows of constant depth), so that every (t0 / depth, px / width) division -- or, under a rotated
    calibration, every guard decision -- the tap chain can meet at this geometry is compared with the oracle."""
    kw = dict(SWEEPS[sweep])
    cw, ch = kw.pop("cw", w), kw.pop("ch", h)
    cal, desc = calib_and_desc(w, h, cw, ch, tf=synth.TF_STITCH[5], **kw)
    ctx.set_stream(0, desc)
    n_frames = -(-65536 // h)
    col = synth.color_frame(cw, ch, 11, 0)
    jobs = []
    for f in range(n_frames):
        z = ((np.arange(h, dtype=np.int64) + f * h) % 65536).astype(np.uint16)
        jobs.append((0, np.repeat(z[:, None], w, axis=1), col))
    for lo in range(0, n_frames, 32):
        chunk = jobs[lo:lo + 32]
        got = run_batch(ctx, chunk, None)
        for (_, z, c), (rec, _, _) in zip(chunk, got):
            want = R.frame(cal, z, c, 3, cw * 3, synth.TF_STITCH[5])
            bad = np.nonzero((rec != want).any(axis=1))[0]
            assert bad.size == 0, "z16=%d x=%d: got %s want %s" % (
                z.reshape(-1)[bad[0]], bad[0] % w, rec[bad[0]], want[bad[0]])


@pytest.mark.parametrize("geom", ["720p_color1080p", "480p_color720p_odd", "480p_color720p_no_map"])
def test_every_depth_value_at_every_row_with_a_row_map(ctx, R, geom):
    """Colour of another size behind a pure x baseline runs row-exact through a host-built row map (pcs_abi.cu
    make_rowmap): the tap row must not depend on the depth value.  Exhaustive: every z16 in 1..65535 at EVERY row
    (column x of frame f carries depth f * w + x), against the oracle's full chain."""
    if geom == "720p_color1080p":       # what the reference records (src/pcs-camera-grab-frames.cpp:69-70)
        w, h, kw = 1280, 720, dict(cw=1920, ch=1080)
    elif geom == "480p_color720p_odd":  # non-round vertical intrinsics whose rows stay 0.17 px away from a tap boundary
        w, h, kw = 848, 480, dict(cw=1280, ch=720, dfy=421.3, dppy=236.2, cfy=631.9, cppy=361.1)
    else:                               # a row comes within 1e-3 px of a boundary: no map, the windowed kernel runs
        w, h, kw = 848, 480, dict(cw=1280, ch=720, dfy=421.3, dppy=236.2, cfy=633.1, cppy=361.4)
    cw, ch = kw.pop("cw"), kw.pop("ch")
    cal, desc = calib_and_desc(w, h, cw, ch, tf=synth.TF_STITCH[2], translation=synth.D2C_BASELINE, **kw)
    ctx.set_stream(0, desc)
    col = synth.color_frame(cw, ch, 12, 0)
    n_frames = -(-65536 // w)
    jobs = []
    for f in range(n_frames):
        z = ((np.arange(w, dtype=np.int64) + f * w) % 65536).astype(np.uint16)
        jobs.append((0, np.ascontiguousarray(np.repeat(z[None, :], h, axis=0)), col))
    for lo in range(0, n_frames, 16):
        chunk = jobs[lo:lo + 16]
        got = run_batch(ctx, chunk, None)
        for (_, z, c), (rec, _, _) in zip(chunk, got):
            want = R.frame(cal, z, c, 3, cw * 3, synth.TF_STITCH[2])
            bad = np.nonzero((rec != want).any(axis=1))[0]
            assert bad.size == 0, "z16=%d row=%d: got %s want %s" % (
                z.reshape(-1)[bad[0]], bad[0] // w, rec[bad[0]], want[bad[0]])


def test_heterogeneous_batch(ctx, R):
    # several streams of different geometry and tex mode in one batch
    specs = [dict(w=1280, h=720), dict(w=848, h=480, translation=synth.D2C_BASELINE),
             dict(w=640, h=480, translation=(0.01, 0, 0),
                  rotation=None if ctx.variant == 2 else small_rotation(0.02, 0.01, -0.01))]
    cals, jobs = [], []
    for s, kw in enumerate(specs):
        kw = dict(kw)
        w, h = kw.pop("w"), kw.pop("h")
        cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[s], **kw)
        ctx.set_stream(s, desc)
        cals.append((cal, w, h))
        for f in range(3):
            jobs.append((s, synth.depth_frame(w, h, s, f), synth.color_frame(w, h, s, f)))
    got = run_batch(ctx, jobs, None)
    for (s, z, col), (rec, _, _) in zip(jobs, got):
        cal, w, h = cals[s]
        assert np.array_equal(rec, R.frame(cal, z, col, 3, w * 3, synth.TF_STITCH[s]))


def test_float_xyzrgb_output(ctx0, R):
    w, h = 640, 360
    cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[4], translation=synth.D2C_BASELINE)
    ctx0.set_stream(0, desc)
    z, col = synth.depth_frame(w, h, 1, 2), synth.color_frame(w, h, 1, 2)
    (rec, fo, _), = run_batch(ctx0, [(0, z, col)], None, float_out=True)
    xyz, uv = R.deproject(cal, z)
    want_rec = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_STITCH[4])
    assert np.array_equal(rec, want_rec)
    want_xyz = R.transform_points(xyz, synth.TF_STITCH[4])
    # north_star tolerance: 1e-5 relative on float XYZ (the kernel's FMA chain is in fact exact)
    assert np.allclose(fo[:, :3], want_xyz, rtol=1e-5, atol=0)
    bgra = fo[:, 3].copy().view(np.uint32)
    r, g_, b = want_rec[:, 3].view(np.uint16) & 0xFF, want_rec[:, 3].view(np.uint16) >> 8, want_rec[:, 4] & 0xFF
    assert np.array_equal(bgra, (0xFF << 24) | (r.astype(np.uint32) << 16) | (g_.astype(np.uint32) << 8) | b.astype(np.uint32))


@pytest.mark.parametrize("lane_reversed", [True, False])
def test_cutoff_compaction(ctx0, R, lane_reversed):
    w, h = 1280, 720
    cal, desc = calib_and_desc(w, h, tf=synth.TF_CAMERA, translation=synth.D2C_BASELINE, cutoff=True)
    desc.cutoff_lane_reversed = int(lane_reversed)
    ctx0.set_stream(3, desc)
    z = synth.depth_frame(w, h, 4, 4, lo=300, hi=2600)
    col = synth.color_frame(w, h, 4, 4)
    xyz, uv = R.deproject(cal, z)
    if lane_reversed:
        want = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, cutoff=True)   # reference order + quirk
    else:
        dense = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA)
        keep = (xyz[:, 2] > 0) & (xyz[:, 2] <= 1.5) & (xyz[:, 0] > -2) & (xyz[:, 0] <= 2)
        want = dense[keep]
    assert 1000 < len(want) < w * h
    (rec, _, cnt), = run_batch(ctx0, [(3, z, col)], None, want_count=True)
    assert cnt == len(want) and np.array_equal(rec, want)
    size, buf = ctx0.send_xyzrgb(3, z, col, write_header=True)
    assert size == len(want) * 10 and buf.view(np.int32)[0] == size
    assert np.array_equal(buf[2:2 + want.size].reshape(-1, 5), want)
    assert np.all(buf[2 + want.size:2500000] == 0)              # memset region past the records
    assert np.all(buf[2500000:] == 0x5A5A)                      # nothing beyond BUF_SIZE bytes is touched
    got = ctx0.pack_from_vertices(3, xyz, uv, col)
    assert np.array_equal(got, want)


def test_from_vertices_full_frame_vs_oracle(ctx0, R):
    w, h = 1280, 720
    cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[2], translation=synth.D2C_BASELINE)
    ctx0.set_stream(0, desc)
    z, col = synth.depth_frame(w, h, 5, 1), synth.color_frame(w, h, 5, 1)
    xyz, uv = R.deproject(cal, z)
    got = ctx0.pack_from_vertices(0, xyz, uv, col)
    assert np.array_equal(got, R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_STITCH[2]))
    # n % 8 == 4 tail, and n = 0
    assert np.array_equal(ctx0.pack_from_vertices(0, xyz[:1004], uv[:1004], col),
                          R.pack(xyz[:1004], uv[:1004], col, w, h, 3, w * 3, synth.TF_STITCH[2]))
    assert len(ctx0.pack_from_vertices(0, xyz[:0], uv[:0], col)) == 0
    with pytest.raises(pcs.PcsError):
        ctx0.pack_from_vertices(0, xyz[:6], uv[:6], col)        # reference needs n % 4 == 0


def test_linearity_of_transform_property(ctx0):
    # size-independent property at full size: records of frame under tf2 o identity-depth shift...
    # translating the world transform by whole millimetres shifts every int16 coordinate by that much
    w, h = 1280, 720
    z, col = synth.depth_frame(w, h, 6, 0), synth.color_frame(w, h, 6, 0)
    tf_a = synth.IDENTITY.copy()
    tf_b = synth.IDENTITY.copy()
    tf_b[3], tf_b[7], tf_b[11] = 1.0, -2.0, 3.0                 # exact in fp32, result exact multiples
    ctx0.set_stream(0, pcs.stream_desc(w, h, tf=tf_a))
    ctx0.set_stream(1, pcs.stream_desc(w, h, tf=tf_b))
    (a, _, _), (b, _, _) = run_batch(ctx0, [(0, z, col), (1, z, col)], None)
    assert np.array_equal(a[:, 3:], b[:, 3:])
    d = b[:, :3].astype(np.int32) - a[:, :3].astype(np.int32)
    # |coords| < 8 m here, so x + 1.0 etc. round at 2^-21 and truncation is toward zero: the
    # difference is the shift to within one LSB, and exactly the shift for almost every z (> 0)
    assert np.all(np.abs(d - np.array([1000, -2000, 3000])) <= 1)
    assert (d[:, 2] == 3000).mean() > 0.99


def test_async_begin_end_many_streams(ctx0, R):
    w, h = 848, 480
    bufs, want = [], []
    for s in range(4):
        cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[s], translation=synth.D2C_BASELINE)
        ctx0.set_stream(s, desc)
        z = ctx0.host_alloc(w * h * 2, np.uint16)
        z[:] = synth.depth_frame(w, h, s, 7).reshape(-1)
        col = ctx0.host_alloc(w * h * 3, np.uint8)
        col[:] = synth.color_frame(w, h, s, 7).reshape(-1)
        buf = ctx0.new_camera_buffer(pinned=True)
        bufs.append((z, col, buf))
        want.append(R.frame(cal, z.reshape(h, w), col.reshape(h, w * 3), 3, w * 3, synth.TF_STITCH[s]))
    for s, (z, col, buf) in enumerate(bufs):
        ctx0.send_begin(s, z, col, buf, write_header=True)
    with pytest.raises(pcs.PcsError):
        ctx0.send_begin(0, *bufs[0])                            # one frame in flight per stream
    for s, (z, col, buf) in enumerate(bufs):
        assert ctx0.send_end(s) == w * h * 10
        assert np.array_equal(buf[2:2 + w * h * 5].reshape(-1, 5), want[s])
        assert buf.view(np.int32)[0] == w * h * 10


def test_error_paths(ctx0):
    with pytest.raises(pcs.PcsError) as e:
        ctx0.send_xyzrgb(7, np.zeros(8, np.uint16), np.zeros(24, np.uint8))
    assert e.value.status == pcs.PCS_ERR_INVALID and "not configured" in str(e.value)
    with pytest.raises(pcs.PcsError):
        ctx0.set_stream(99, pcs.stream_desc(8, 1))
    with pytest.raises(pcs.PcsError):
        ctx0.set_stream(0, pcs.stream_desc(8, 1, bpp=2))
    # rs2_distortion models librealsense does not apply on this side of the chain are refused, not ignored
    bad = pcs.stream_desc(64, 4, color_distortion=(0.1, 0, 0, 0, 0))
    bad.color.model = 3                                           # F-Theta
    with pytest.raises(pcs.PcsError) as e:
        ctx0.set_stream(5, bad)
    assert e.value.status == pcs.PCS_ERR_UNSUPPORTED
    bad = pcs.stream_desc(64, 4, depth_distortion=(0.1, 0, 0, 0, 0))
    bad.depth.model = 5                                           # Kannala-Brandt
    with pytest.raises(pcs.PcsError) as e:
        ctx0.set_stream(5, bad)
    assert e.value.status == pcs.PCS_ERR_UNSUPPORTED
    ctx0.set_stream(5, pcs.stream_desc(1284, 4))                  # width % 8 != 0
    with pytest.raises(pcs.PcsError) as e:
        ctx0.send_xyzrgb(5, np.zeros(1284 * 4, np.uint16), np.zeros(1284 * 4 * 3, np.uint8))
    assert e.value.status == pcs.PCS_ERR_UNSUPPORTED
    ctx0.set_stream(5, pcs.stream_desc(1280, 720))
    t = torch.zeros(1280 * 720 + 8, dtype=torch.int16, device="cuda")
    with pytest.raises(pcs.PcsError):                            # misaligned depth pointer
        ctx0.batch([(5, t.data_ptr() + 2, t.data_ptr(), t.data_ptr())])


# ------------------------------------------------------------------ guarded taps: edge geometries
ROT_EDGE = {
    # a frame lower than a segment window is tall, segments of the last partial 128 px, a single row
    "rot_256x8": dict(w=256, h=8, rotation=small_rotation(0.002, -0.001, 0.004), translation=(0.015, 0.0002, 0.0)),
    "rot_200x1": dict(w=200, h=1, rotation=small_rotation(0.001, 0.001, -0.002), translation=(0.012, 0.0, 0.0001)),
    # 2048-px colour (16 segments, the most the table holds) behind a 1024-px depth frame, padded colour stride
    "rot_1024x64_color2048": dict(w=1024, h=64, cw=2048, ch=128, stride=2048 * 3 + 64,
                                  rotation=small_rotation(-0.002, 0.002, 0.003), translation=(0.02, 0.0003, -0.0002)),
    # colour SMALLER than depth: several depth rows tap the same colour row
    "rot_848x480_color424": dict(w=848, h=480, cw=424, ch=240, rotation=small_rotation(0.003, 0.002, -0.004),
                                 translation=(0.015, 0.0, 0.0)),
    # negative T.z and a strong x baseline: t2 shrinks for near points, the guard depth handles them
    "rot_640x360_tz_negative": dict(w=640, h=360, rotation=small_rotation(0.004, 0.004, 0.004), translation=(0.05, 0.001, -0.004)),
}


@pytest.mark.parametrize("case", list(ROT_EDGE))
def test_guarded_taps_edge_geometries(ctx, R, case):
    kw = dict(ROT_EDGE[case])
    w, h = kw.pop("w"), kw.pop("h")
    cw, ch = kw.pop("cw", w), kw.pop("ch", h)
    stride = kw.pop("stride", cw * 3)
    cal, desc = calib_and_desc(w, h, cw, ch, tf=synth.TF_STITCH[4], stride=stride, **kw)
    ctx.set_stream(0, desc)
    jobs = []
    for f in range(3):
        # very near depths (under the guard depth), far ones and holes in one frame
        z = synth.depth_frame(w, h, 5, f, lo=1 if f == 0 else 300, hi=900 if f == 0 else 6000)
        jobs.append((0, z, synth.color_frame(cw, ch, 5, f, stride=stride)))
    try:
        got = run_batch(ctx, jobs, None)
    except pcs.PcsError as e:       # kernel_variant 2 may refuse a geometry, never return wrong bytes
        assert ctx.variant == 2 and e.status == pcs.PCS_ERR_UNSUPPORTED
        return
    for (_, z, col), (rec, _, _) in zip(jobs, got):
        want = R.frame(cal, z, col, 3, stride, synth.TF_STITCH[4])
        bad = np.nonzero((rec != want).any(axis=1))[0]
        assert bad.size == 0, "first mismatches at points %s: got %s want %s" % (bad[:5], rec[bad[:5]], want[bad[:5]])


def test_transform_changed_under_a_live_batch_and_many_streams(R):
    """The camera->world transforms of a pipelined launch travel in the kernel parameters (64 slots per launch): 70
    streams with 70 different transforms split the batch into two launches, and a transform set AFTER the batch was
    created must be the one the next run uses (include/pcs_b200.h: only tf may change under a live batch)."""
    w, h, n_streams = 256, 16, 70
    ctx = pcs.Context(device=0, max_streams=n_streams, kernel_variant=2)
    rng = np.random.default_rng(5)

    def tf_of(k):
        t = synth.TF_STITCH[k % 8].copy()
        t[3], t[7], t[11] = 0.01 * k, -0.02 * k, 0.005 * k
        return t

    for rot in (None, small_rotation(0.002, 0.001, -0.003)):
        cal = None
        for k in range(n_streams):
            cal, desc = calib_and_desc(w, h, tf=tf_of(k), translation=synth.D2C_BASELINE, rotation=rot)
            ctx.set_stream(k, desc)
        z = synth.depth_frame(w, h, 1, 1)
        col = synth.color_frame(w, h, 1, 1)
        dz, dc = torch.from_numpy(z.view(np.int16)).cuda(), torch.from_numpy(col).cuda()
        pays = [torch.zeros(w * h * 5, dtype=torch.int16, device="cuda") for _ in range(n_streams)]
        b = ctx.batch([(k, dz.data_ptr(), dc.data_ptr(), pays[k].data_ptr()) for k in range(n_streams)])
        assert b.launches == 2
        b.run(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for k in (0, 1, 63, 64, 69):
            assert np.array_equal(pays[k].cpu().numpy().reshape(-1, 5), R.frame(cal, z, col, 3, w * 3, tf_of(k)))
        # new transforms for two streams, one in each launch; same batch
        for k in (3, 66):
            _, desc = calib_and_desc(w, h, tf=tf_of(k + 100), translation=synth.D2C_BASELINE, rotation=rot)
            ctx.set_stream(k, desc)
        b.run(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for k in (3, 66):
            assert np.array_equal(pays[k].cpu().numpy().reshape(-1, 5), R.frame(cal, z, col, 3, w * 3, tf_of(k + 100)))
        assert np.array_equal(pays[4].cpu().numpy().reshape(-1, 5), R.frame(cal, z, col, 3, w * 3, tf_of(4)))
        b.close()
    del rng


def test_distortion_on_the_side_that_does_not_apply_it_has_no_effect(ctx, R):
    """rsutil.h undistorts only when deprojecting (inverse Brown-Conrady) and distorts only when projecting (modified
    Brown-Conrady): a D455 colour stream reports inverse Brown-Conrady coefficients, which rs2_project_point_to_pixel
    ignores.  Such a stream is accepted, keeps the fast kernels and gives the bytes of the undistorted calibration."""
    w, h = 848, 480
    cal, plain = calib_and_desc(w, h, tf=synth.TF_STITCH[2], translation=synth.D2C_BASELINE)
    _, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[2], translation=synth.D2C_BASELINE)
    desc.color.model, desc.color.coeffs = 2, (C.c_float * 5)(-0.05, 0.06, 0.0003, -0.0004, -0.02)      # inverse BC on colour
    desc.depth.model, desc.depth.coeffs = 4, (C.c_float * 5)(0.1, 0.1, 0.0, 0.0, 0.0)                   # plain BC on depth
    ctx.set_stream(0, desc)
    z, col = synth.depth_frame(w, h, 4, 4), synth.color_frame(w, h, 4, 4)
    (rec, _, _), = run_batch(ctx, [(0, z, col)], None)       # (kernel_variant 2 accepts it: still the x-baseline mode)
    assert np.array_equal(rec, R.frame(cal, z, col, 3, w * 3, synth.TF_STITCH[2]))
    del plain


def test_cutoff_one_pass_edge_cases(ctx0, R):
    """The one-pass -c of k1_direct (persistent blocks over ticketed tiles, look-back, compaction in shared memory): a batch
    of several frames whose last tile is partial (848x480 = 198.75 tiles), a box that keeps every valid point, one that
    keeps none, the reference's box with its reversed lanes, a rotated calibration, and payloads at every 2-byte
    alignment inside their 16 bytes."""
    w, h = 848, 480
    n = w * h
    boxes = {   # stream -> (lane_reversed, z_lo, z_hi, x_lo, x_hi, calibration kw)
        0: (True, 0.0, 1.5, -2.0, 2.0, dict(translation=synth.D2C_BASELINE)),                   # the reference's -c
        1: (False, 0.0, 100.0, -100.0, 100.0, dict(translation=synth.D2C_BASELINE)),            # every valid point
        2: (False, 0.0, 1e-4, -2.0, 2.0, dict()),                                               # nothing
        3: (False, 0.5, 2.5, -0.4, 1.1, dict(translation=(0.0149, 0.0002, -0.0003), rotation=small_rotation(0.003, -0.002, 0.004))),
    }
    cals = {}
    for s, (rev, z_lo, z_hi, x_lo, x_hi, kw) in boxes.items():
        cal, desc = calib_and_desc(w, h, tf=synth.TF_STITCH[s], cutoff=True, **kw)
        desc.cutoff_lane_reversed, desc.z_lo, desc.z_hi, desc.x_lo, desc.x_hi = int(rev), z_lo, z_hi, x_lo, x_hi
        ctx0.set_stream(s, desc)
        cals[s] = cal
    keep, jobs = [], []
    for f in range(10):
        s = f % 4
        z, col = synth.depth_frame(w, h, 20 + s, f, lo=300, hi=4000), synth.color_frame(w, h, 20 + s, f)
        dz, dc = dev(z), dev(col)
        pay = torch.full((n * 5 + 16,), 0x5A5A, dtype=torch.int16, device="cuda")
        cnt = torch.full((1,), -1, dtype=torch.int32, device="cuda")
        shift = f % 8                                              # records start 2 * shift bytes into the allocation
        keep.append((s, z, col, dz, dc, pay, cnt, shift))
        jobs.append((s, dz.data_ptr(), dc.data_ptr(), pay.data_ptr() + 2 * shift, None, cnt.data_ptr()))
    b = ctx0.batch(jobs)
    assert b.launches <= 4                                         # one launch per group of frames, none for the compaction
    b.run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for s, z, col, _, _, pay, cnt, shift in keep:
        rev, z_lo, z_hi, x_lo, x_hi, _ = boxes[s]
        xyz, uv = R.deproject(cals[s], z)
        if rev:
            want = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_STITCH[s], cutoff=True)
        else:
            dense = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_STITCH[s])
            m = (xyz[:, 2] > np.float32(z_lo)) & (xyz[:, 2] <= np.float32(z_hi)) & (xyz[:, 0] > np.float32(x_lo)) & (xyz[:, 0] <= np.float32(x_hi))
            want = dense[m]
        c = int(cnt.item())
        assert c == len(want), (s, c, len(want))
        got = pay.cpu().numpy()
        assert np.array_equal(got[shift:shift + c * 5].reshape(-1, 5), want), s
        assert np.all(got[:shift] == 0x5A5A) and np.all(got[shift + c * 5:] == 0x5A5A), "wrote outside the kept records"
    assert int(keep[2][6].item()) == 0 and int(keep[1][6].item()) == int((keep[1][1] != 0).sum())
    b.close()
