"""CPU tests: pin oracle/pcs_oracle.c (the restatement) against

* the golden fixtures produced by the reference's own compiled functions
  (tests/golden/make_golden.py) -- always;
* the compiled reference itself on larger seeded inputs -- when oracle/_ref is
  present (dev container, or the prebuilt files on the GPU box).
"""
import numpy as np
import pytest

from conftest import load_golden, roundtrip_inputs
from pointcloud_stitching_b200 import synth


# ---------------------------------------------------------------- golden: camera side
@pytest.mark.parametrize("name", ["pack_96x64_identity", "pack_96x64_baseline", "pack_96x64_cutoff"])
def test_pack_golden(restatement, name):
    g = load_golden(name)
    w, h = int(g["w"]), int(g["h"])
    out = restatement.pack(g["xyz"], g["uv"], g["color"], w, h, 3, w * 3, g["tf"], bool(g["cutoff"]))
    assert np.array_equal(out, g["out_records"])


@pytest.mark.parametrize("name", ["pack_96x64_identity", "pack_96x64_baseline", "pack_96x64_cutoff"])
def test_deproject_regenerates_golden_inputs(restatement, name):
    # the fixture's vertices came out of the deprojection spec: keep that reproducible
    import oracle
    g = load_golden(name)
    cal = oracle.make_calib(int(g["w"]), int(g["h"]), translation=tuple(g["translation"]))
    xyz, uv = restatement.deproject(cal, g["z16"])
    assert xyz.tobytes() == g["xyz"].tobytes() and uv.tobytes() == g["uv"].tobytes()


@pytest.mark.parametrize("name", ["pack_adversarial", "pack_adversarial_cutoff"])
def test_pack_adversarial_golden(restatement, name):
    g = load_golden(name)
    out = restatement.pack(g["xyz"], g["uv"], g["color"], int(g["w"]), int(g["h"]), 3,
                           int(g["stride"]), g["tf"], bool(g["cutoff"]))
    assert np.array_equal(out, g["out_records"])


def test_send_golden(restatement):
    g = load_golden("send_64x32")
    w, h = int(g["w"]), int(g["h"])
    n5 = w * h * 5 + 64
    size, buf = restatement.send(g["xyz"], g["uv"], g["color"], w, h, 3, w * 3, g["tf"])
    assert size == int(g["size"]) == w * h * 10
    assert np.array_equal(buf[:n5], g["out_head_nosend"])          # header stays 0 without -s
    assert np.array_equal(buf[2499990:2500010], g["out_memset_edge"])  # memset stops at byte 5e6
    size, buf = restatement.send(g["xyz"], g["uv"], g["color"], w, h, 3, w * 3, g["tf"],
                                 write_header=True)
    assert np.array_equal(buf[:n5], g["out_head_send"])
    assert np.array_equal(buf.view(np.uint8)[: size + 4], g["wire_bytes"])
    assert buf.view(np.int32)[0] == size


# ---------------------------------------------------------------- golden: stitch side
def test_stitch_golden(restatement):
    g = load_golden("stitch_2048")
    rec = g["records"]
    for d in (1, 2, 3, 4, 7):
        assert np.array_equal(restatement.concat([rec], d), g["out_raw_d%d" % d])
    for d in (1, 2, 4):
        assert restatement.unpack(rec, d).tobytes() == g["out_unpack_d%d" % d].tobytes()
        for k in (0, 5):
            got = restatement.pcl_stitch([rec], [synth.TF_STITCH[k]], d)
            assert np.array_equal(got, g["out_pcl_d%d_tf%d" % (d, k)])


def test_roundtrip_golden(restatement):
    # int16 -> /1000.0f -> *1000.0f -> int16 is lossy (740 values)
    allv = roundtrip_inputs()
    diff = load_golden("roundtrip_all_int16")["out_minus_in"]
    rt = restatement.repack(restatement.unpack(allv))
    assert np.array_equal(rt.astype(np.int32) - allv.astype(np.int32), diff.astype(np.int32))
    assert int((diff[:, 0] != 0).sum()) == 740


# ---------------------------------------------------------------- live reference
@pytest.mark.parametrize("w,h", [(1280, 720), (848, 480)])
@pytest.mark.parametrize("trans", [(0, 0, 0), synth.D2C_BASELINE])
def test_pack_vs_compiled_reference(restatement, ref_camera, w, h, trans):
    import oracle
    cal = oracle.make_calib(w, h, translation=trans)
    z, col = synth.depth_frame(w, h, 2, 5), synth.color_frame(w, h, 2, 5)
    xyz, uv = restatement.deproject(cal, z)
    for cutoff in (False, True):
        a = restatement.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, cutoff)
        b = ref_camera.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, cutoff, threads=1)
        assert np.array_equal(a, b)
    # without -c the thread count does not change the bytes
    b4 = ref_camera.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, threads=4)
    assert np.array_equal(restatement.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA), b4)


def test_scalar_path_is_not_the_oracle(restatement, ref_camera):
    # the scalar loop differs from the SIMD loop by +-1 LSB on some coordinates
    import oracle
    w, h = 320, 240
    cal = oracle.make_calib(w, h)
    z, col = synth.depth_frame(w, h), synth.color_frame(w, h)
    xyz, uv = restatement.deproject(cal, z)
    simd = ref_camera.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, simd=True)
    scal = ref_camera.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, simd=False)
    d = simd.astype(np.int32) - scal.astype(np.int32)
    assert np.array_equal(simd[:, 3:], scal[:, 3:])
    assert 0 < int((d != 0).sum()) and int(np.abs(d).max()) == 1


def test_stitch_vs_compiled_reference(restatement, ref_client, ref_optimized):
    rng = np.random.default_rng(7)
    rec = rng.integers(-32768, 32768, (8000, 5), dtype=np.int16)
    rec[:, 4] &= 0xFF
    for d in (1, 2, 5):
        assert np.array_equal(restatement.concat([rec], d), ref_client.raw_stitch_1cam(rec, d))
        assert restatement.unpack(rec, d).tobytes() == ref_optimized.unpack(rec, d).tobytes()
        for k in range(8):
            a = restatement.pcl_stitch([rec], [synth.TF_STITCH[k]], d)
            assert np.array_equal(a, ref_optimized.pcl_stitch_1cam(rec, synth.TF_STITCH[k], d))
            assert np.array_equal(a, ref_client.pcl_stitch_1cam(rec, synth.TF_STITCH[k], d))


def test_reference_replay_runs(ref_camera):
    # the reference's own main() replay loop over synthetic frames prints its summary
    import oracle
    w, h = 320, 240
    d = np.stack([synth.depth_frame(w, h, 0, f) for f in range(3)])
    c = np.stack([synth.color_frame(w, h, 0, f) for f in range(3)])
    avg, calc, log = ref_camera.replay(oracle.make_calib(w, h), d, c, 3, w * 3, synth.TF_CAMERA)
    assert avg > 0 and calc > 0
    assert "### Total Frames = 3" in log and "# Points : %d" % (w * h) in log


# ---------------------------------------------------------------- spec-only pieces
def test_deproject_spec_properties(restatement):
    import oracle
    w, h = 256, 144
    z = synth.depth_frame(w, h, 1, 1)
    xyz, uv = restatement.deproject(oracle.make_calib(w, h), z)
    hole = z.reshape(-1) == 0
    assert hole.any()
    assert np.all(uv[hole] == 0) and np.all(xyz[hole, 2] == 0)
    # SPEC.md s1: aligned streams sample their own pixel
    xi = np.trunc(uv[:, 0] * np.float32(w) + np.float32(0.5)).astype(int)
    yi = np.trunc(uv[:, 1] * np.float32(h) + np.float32(0.5)).astype(int)
    px, py = np.tile(np.arange(w), h), np.repeat(np.arange(h), w)
    assert np.array_equal(xi[~hole], px[~hole]) and np.array_equal(yi[~hole], py[~hole])
    # a 15 mm baseline shifts taps right by fx*0.015/z pixels
    xyz2, uv2 = restatement.deproject(oracle.make_calib(w, h, translation=synth.D2C_BASELINE), z)
    xi2 = np.trunc(uv2[:, 0] * np.float32(w) + np.float32(0.5)).astype(int)
    expect = px + (w / 2) * 0.015 / np.maximum(xyz[:, 2], 1e-9)
    assert np.all(np.abs(xi2[~hole] - expect[~hole]) <= 1.0)
    assert xyz.tobytes() == xyz2.tobytes()


def _np_deproject(cal, z16):
    """oracle/SPEC.md s1 written again in numpy float32 (every operator rounds once): an independent reading of the
    same text, distortion included."""
    f = np.float32
    W, H = cal.depth.width, cal.depth.height
    x, y = np.meshgrid(np.arange(W, dtype=f), np.arange(H, dtype=f))
    nx, ny = (x - f(cal.depth.ppx)) / f(cal.depth.fx), (y - f(cal.depth.ppy)) / f(cal.depth.fy)

    def radial(k, r2):
        return f(1) + k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2

    if cal.depth.model == 2:
        k = [f(v) for v in cal.depth.coeffs]
        r2 = nx * nx + ny * ny
        fr = radial(k, r2)
        ux = nx * fr + f(2) * k[2] * nx * ny + k[3] * (r2 + f(2) * nx * nx)
        uy = ny * fr + f(2) * k[3] * nx * ny + k[2] * (r2 + f(2) * ny * ny)
        nx, ny = ux, uy
    depth = f(cal.depth_scale) * z16.reshape(H, W).astype(f)
    p0, p1, p2 = depth * nx, depth * ny, depth
    R, T = [f(v) for v in cal.rotation], [f(v) for v in cal.translation]
    with np.errstate(all="ignore"):
        t0 = R[0] * p0 + R[3] * p1 + R[6] * p2 + T[0]
        t1 = R[1] * p0 + R[4] * p1 + R[7] * p2 + T[1]
        t2 = R[2] * p0 + R[5] * p1 + R[8] * p2 + T[2]
        qx, qy = t0 / t2, t1 / t2
        if cal.color.model == 1:
            k = [f(v) for v in cal.color.coeffs]
            r2 = qx * qx + qy * qy
            fr = radial(k, r2)
            qx, qy = qx * fr, qy * fr
            dx = qx + f(2) * k[2] * qx * qy + k[3] * (r2 + f(2) * qx * qx)
            dy = qy + f(2) * k[3] * qx * qy + k[2] * (r2 + f(2) * qy * qy)
            qx, qy = dx, dy
        u = (qx * f(cal.color.fx) + f(cal.color.ppx)) / f(cal.color.width)
        v = (qy * f(cal.color.fy) + f(cal.color.ppy)) / f(cal.color.height)
    hole = p2 == 0
    u[hole], v[hole] = 0, 0
    return np.stack([p0, p1, p2], -1).reshape(-1, 3), np.stack([u, v], -1).reshape(-1, 2)


@pytest.mark.parametrize("kw", [
    dict(),
    dict(translation=synth.D2C_BASELINE, color_distortion=(0.12, -0.25, 0.0012, -0.0008, 0.09)),
    dict(depth_distortion=(-0.05, 0.07, 0.0005, 0.0011, -0.02)),
    dict(cw=480, ch=270, translation=(0.015, 0.001, -0.002),
         rotation=(0.99998, 0.005, 0.003, -0.005, 0.99998, -0.004, -0.003, 0.004, 0.99998),
         depth_distortion=(0.02, -0.03, 0.0004, -0.0006, 0.01), color_distortion=(0.1, -0.21, 0.001, 0.0015, 0.07)),
])
def test_deproject_distortion_against_a_second_reading_of_the_spec(restatement, kw):
    import oracle
    w, h = 320, 180
    cal = oracle.make_calib(w, h, **kw)
    z = synth.depth_frame(w, h, 3, 2)
    xyz, uv = restatement.deproject(cal, z)
    xyz2, uv2 = _np_deproject(cal, z)
    assert xyz.tobytes() == xyz2.tobytes() and uv.tobytes() == uv2.tobytes()


def test_deproject_distortion_properties(restatement):
    import oracle
    w, h = 320, 180
    z = synth.depth_frame(w, h, 3, 2)
    plain = restatement.deproject(oracle.make_calib(w, h, translation=synth.D2C_BASELINE), z)
    # all-zero coefficients: the radial factor is exactly 1 and the tangential terms exactly 0
    zero = restatement.deproject(oracle.make_calib(w, h, translation=synth.D2C_BASELINE, depth_distortion=(0,) * 5,
                                                   color_distortion=(0,) * 5), z)
    assert np.array_equal(plain[0], zero[0]) and np.array_equal(plain[1], zero[1])
    # a barrel colour lens pulls taps towards the principal point, more so far from it; the vertices do not move
    bar = restatement.deproject(oracle.make_calib(w, h, translation=synth.D2C_BASELINE,
                                                  color_distortion=(-0.2, 0, 0, 0, 0)), z)
    assert plain[0].tobytes() == bar[0].tobytes()
    valid = z.reshape(-1) != 0
    c = np.array([(w - 1) / 2 / w, (h - 1) / 2 / h], np.float32)
    r_plain = np.linalg.norm((plain[1] - c) * (w, h), axis=1)[valid]
    r_bar = np.linalg.norm((bar[1] - c) * (w, h), axis=1)[valid]
    far = r_plain > 20
    assert np.all(r_bar[far] < r_plain[far])
    shrink = 1 - r_bar[far] / r_plain[far]
    assert np.corrcoef(shrink, r_plain[far] ** 2)[0, 1] > 0.99


def test_voxel_merge_spec(restatement):
    rec = np.array([[5, 5, 5, 0x0201, 3], [9, 0, 1, 0x0403, 6], [10, 0, 0, 0x1010, 0x10],
                    [-1, -10, -11, 0xFF, 0xFF], [-10, -1, -20, 0x01, 0x01]], np.int16)
    out = restatement.voxel_merge(rec, 10)
    # voxels: (-1,-1,-2) <- pts 3,4 ; (0,0,0) <- pts 0,1 ; (1,0,0) <- pt 2 ; ascending (kz,ky,kx)
    assert out.tolist() == [
        [-10 + (9 + 0) // 2, -10 + (0 + 9) // 2, -20 + (9 + 0) // 2, (0xFF + 1) // 2, (0xFF + 1) // 2],
        [(5 + 9) // 2, (5 + 0) // 2, (5 + 1) // 2, ((2 + 4) // 2 << 8) | (1 + 3) // 2, (3 + 6) // 2],
        [10, 0, 0, 0x1010, 0x10]]
    rng = np.random.default_rng(3)
    big = rng.integers(-2000, 2000, (50000, 5)).astype(np.int16)
    big[:, 3] = rng.integers(0, 65536, 50000).astype(np.uint16).view(np.int16)
    big[:, 4] = rng.integers(0, 256, 50000)
    m = restatement.voxel_merge(big, 10)
    k = np.floor_divide(m[:, :3].astype(np.int32), 10)
    key = (k[:, 2].astype(np.int64) << 40) + (k[:, 1].astype(np.int64) << 20) + k[:, 0]
    assert np.all(np.diff(key) > 0)                                   # sorted, unique
    assert np.array_equal(restatement.voxel_merge(m, 10), m)          # idempotent
    assert len(m) == len(np.unique(np.floor_divide(big[:, :3].astype(np.int32), 10), axis=0))
    assert np.array_equal(restatement.voxel_merge(big[::-1], 10), m)  # order independent
    assert len(restatement.voxel_merge(big[:0], 10)) == 0
