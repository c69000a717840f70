"""-m gpu parity tests, stitch side (concat / decimate, unpack-transform-append-repack,
voxel merge) through the C ABI against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import load_golden, roundtrip_inputs

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import oracle  # noqa: E402
import pointcloud_stitching_b200 as pcs  # noqa: E402
from gpu_util import dev  # noqa: E402
from pointcloud_stitching_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def R():
    return oracle.restatement()


@pytest.fixture(scope="module")
def ctx():
    c = pcs.Context(device=0, max_streams=2)
    yield c
    c.close()


def random_records(rng, n):
    rec = rng.integers(-32768, 32768, (n, 5), dtype=np.int16)
    rec[:, 4] = rng.integers(0, 256, n)          # the wire format keeps the 10th byte zero
    return rec


def test_stitch_vs_reference_golden(ctx):
    g = load_golden("stitch_2048")
    rec = g["records"]
    for d in (1, 2, 3, 4, 7):
        assert np.array_equal(ctx.stitch_raw([rec], d), g["out_raw_d%d" % d])
    for d in (1, 2, 4):
        for k in (0, 5):
            assert np.array_equal(ctx.stitch_pcl([rec], [synth.TF_STITCH[k]], d), g["out_pcl_d%d_tf%d" % (d, k)])


def test_lossy_roundtrip_vs_reference_golden(ctx):
    # identity transform still changes 740 of the 65 536 int16 values
    allv = roundtrip_inputs()
    diff = load_golden("roundtrip_all_int16")["out_minus_in"].astype(np.int32)
    out = ctx.stitch_pcl([allv], [synth.IDENTITY], 1)
    rec = out[4:].view(np.int16).reshape(-1, 5)
    assert np.array_equal(rec.astype(np.int32) - allv.astype(np.int32), diff)


@pytest.mark.parametrize("n_cams,downsample", [(1, 1), (4, 1), (4, 2), (8, 3), (20, 1), (20, 5)])
def test_multi_camera_vs_oracle(ctx, R, n_cams, downsample):
    rng = np.random.default_rng(n_cams * 10 + downsample)
    sizes = [int(rng.integers(0, 30000)) for _ in range(n_cams)]
    sizes[0] = 40960
    if n_cams > 2:
        sizes[2] = 0                              # a camera that sent nothing
    pay = [random_records(rng, n) for n in sizes]
    tfs = [synth.TF_STITCH[k % 8] for k in range(n_cams)]
    assert np.array_equal(ctx.stitch_raw(pay, downsample), R.concat(pay, downsample))
    assert np.array_equal(ctx.stitch_pcl(pay, tfs, downsample), R.pcl_stitch(pay, tfs, downsample))


@pytest.mark.parametrize("sizes", [[921600, 407040, 8, 256 * 77], [8], [264, 0, 2048]])
def test_vectorised_path_vs_oracle(ctx, R, sizes):
    """No decimation, whole octets per camera, aligned buffers: the warp-tiled 16-byte path
    (stitch_vec) -- every int16 value, every colour byte, against the oracle."""
    rng = np.random.default_rng(len(sizes))
    pay = [random_records(rng, n) for n in sizes]
    pay[0][:65536 if sizes[0] >= 65536 else sizes[0], 0] = np.arange(-32768, 32768)[: min(65536, sizes[0])]
    tfs = [synth.TF_STITCH[(k + 3) % 8] for k in range(len(sizes))]
    d = [dev(p.reshape(-1)) if p.size else torch.zeros(8, dtype=torch.int16, device="cuda") for p in pay]
    total = sum(sizes)
    st = torch.zeros(total * 10 + 32, dtype=torch.uint8, device="cuda")
    cs = torch.cuda.current_stream().cuda_stream
    size = ctx.stitch_raw_dev([t.data_ptr() for t in d], [p.size for p in pay], 1, st.data_ptr() + 12, total * 10 + 4, cs)
    torch.cuda.synchronize()
    assert np.array_equal(st[12:12 + size + 4].cpu().numpy(), R.concat(pay, 1))
    size = ctx.stitch_pcl_dev([t.data_ptr() for t in d], [p.size for p in pay], 1, tfs, st.data_ptr() + 12,
                              total * 10 + 4, None, cs)
    torch.cuda.synchronize()
    assert np.array_equal(st[12:12 + size + 4].cpu().numpy(), R.pcl_stitch(pay, tfs, 1))


def test_device_path_and_cloud32(ctx, R):
    rng = np.random.default_rng(5)
    pay = [random_records(rng, n) for n in (921600, 407040, 8, 123457)]
    tfs = [synth.TF_STITCH[k] for k in range(4)]
    d = [dev(p.reshape(-1)) for p in pay]
    total = sum(p.shape[0] for p in pay)
    st = torch.zeros(total * 10 + 16 + 16, dtype=torch.uint8, device="cuda")
    cloud = torch.zeros(total * 8, dtype=torch.float32, device="cuda")
    base = st.data_ptr() + 12                     # header at +12 so the records are 16-byte aligned
    cs = torch.cuda.current_stream().cuda_stream
    size = ctx.stitch_raw_dev([t.data_ptr() for t in d], [p.size for p in pay], 1, base, total * 10 + 4, cs)
    torch.cuda.synchronize()
    assert np.array_equal(st[12:12 + size + 4].cpu().numpy(), R.concat(pay, 1))
    size = ctx.stitch_pcl_dev([t.data_ptr() for t in d], [p.size for p in pay], 1, tfs, base, total * 10 + 4,
                              cloud.data_ptr(), cs)
    torch.cuda.synchronize()
    assert np.array_equal(st[12:12 + size + 4].cpu().numpy(), R.pcl_stitch(pay, tfs, 1))
    want = np.concatenate([R.transform_cloud(R.unpack(p), t) for p, t in zip(pay, tfs)])
    got = cloud.cpu().numpy().view(oracle.PCLPOINT)
    for f in ("x", "y", "z", "w", "b", "g", "r", "a"):
        assert np.array_equal(got[f], want[f]), f
    # unaligned stitched buffer (header at +0): same bytes through the 2-byte store path
    size = ctx.stitch_raw_dev([t.data_ptr() for t in d], [p.size for p in pay], 2, st.data_ptr(), total * 10 + 4, cs)
    torch.cuda.synchronize()
    assert np.array_equal(st[: size + 4].cpu().numpy(), R.concat(pay, 2))


def test_ply_dump_of_the_stitched_cloud(ctx, R, tmp_path):
    """visualize()'s save path (src/pcs-multicamera-client.cpp:482-489): the stitched PCL cloud as a binary
    PLY -- float x, y, z + uchar red, green, blue per vertex, in stitch order."""
    rng = np.random.default_rng(6)
    pay = [random_records(rng, n) for n in (40000, 8, 12345)]
    tfs = [synth.TF_STITCH[k] for k in range(3)]
    d = [dev(p.reshape(-1)) for p in pay]
    total = sum(p.shape[0] for p in pay)
    st = torch.zeros(total * 10 + 32, dtype=torch.uint8, device="cuda")
    cloud = torch.zeros(total * 8, dtype=torch.float32, device="cuda")
    cs = torch.cuda.current_stream().cuda_stream
    ctx.stitch_pcl_dev([t.data_ptr() for t in d], [p.size for p in pay], 1, tfs, st.data_ptr() + 12, total * 10 + 4,
                       cloud.data_ptr(), cs)
    torch.cuda.synchronize()
    want = np.concatenate([R.transform_cloud(R.unpack(p), t) for p, t in zip(pay, tfs)])
    path = str(tmp_path / "stitched_cloud_0.ply")
    assert ctx.save_ply(cloud.data_ptr(), total, path) == total
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode().split("\n")
    assert header[0] == "ply" and header[1] == "format binary_little_endian 1.0"
    assert "element vertex %d" % total in header and "element camera 1" in header
    props = [h.split()[-1] for h in header[header.index("element vertex %d" % total) + 1:header.index("element camera 1")]]
    assert props == ["x", "y", "z", "red", "green", "blue"]
    row = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    assert len(raw) == end + total * 15 + 21 * 4
    got = np.frombuffer(raw, row, total, end)
    for f in ("x", "y", "z", "r", "g", "b"):
        assert np.array_equal(got[f].view(np.uint32 if f in "xyz" else np.uint8),
                              want[f].view(np.uint32 if f in "xyz" else np.uint8)), f
    # rows on the device: a count that is not a multiple of 32 and an output pointer that is not 16-byte aligned
    rows = torch.zeros(total * 15 + 64, dtype=torch.uint8, device="cuda")
    for n, off in ((total, 0), (total - 7, 1), (33, 5), (1, 3)):
        rows.zero_()
        assert ctx.cloud_to_ply_rows_dev(cloud.data_ptr(), n, rows.data_ptr() + off, cs) == n
        torch.cuda.synchronize()
        h = rows.cpu().numpy()
        assert np.array_equal(np.frombuffer(h[off:off + n * 15].tobytes(), row), got[:n])
        assert not h[off + n * 15:].any() and not h[:off].any()
    assert ctx.save_ply(cloud.data_ptr(), 0, path) == 0


def test_stitch_errors(ctx):
    rec = np.zeros((16, 5), np.int16)
    with pytest.raises(pcs.PcsError):
        ctx.stitch_raw([rec], 0)
    with pytest.raises(pcs.PcsError):
        ctx.stitch_raw([rec] * 33, 1)
    with pytest.raises(pcs.PcsError) as e:
        ctx.stitch_raw_dev([0], [80], 1, 0, 4)
    assert e.value.status == pcs.PCS_ERR_INVALID
    t = torch.zeros(256, dtype=torch.uint8, device="cuda")
    with pytest.raises(pcs.PcsError) as e:       # capacity check
        ctx.stitch_raw_dev([t.data_ptr()], [80], 1, t.data_ptr(), 100)
    assert e.value.status == pcs.PCS_ERR_CAPACITY


@pytest.fixture(scope="module", params=[0, 1, 2, 3, 4], ids=["auto", "pair_sort", "sweep8", "sweep10", "msd"])
def vctx(request):
    """One context per voxel_variant: auto (the one-sweep sort, falling back to the slab partition / the pair sort), the
    (key, index) pair sort, the one-sweep sort (8 / 10 bit), slab partition + bitmap ranking only."""
    c = pcs.Context(device=0, max_streams=2, voxel_variant=request.param)
    c.variant = request.param
    yield c
    c.close()


def test_voxel_merge_vs_oracle(vctx, R):
    rng = np.random.default_rng(11)
    cases = [(50000, 2000, 10), (200000, 400, 10), (5000, 30000, 25), (4, 10, 10), (1, 5, 1), (4096, 300, 10),
             (4097, 300, 10), (8191, 20, 10), (255, 3, 10), (257, 40000, 10), (70001, 60, 3)]
    exact = 0
    for n, span, leaf in cases:
        rec = random_records(rng, n)
        rec[:, :3] = rng.integers(-span, span, (n, 3))
        try:
            got = vctx.voxel_merge(rec, leaf)
        except pcs.PcsError as e:
            # MSD only: a z plane of the occupied box may hold at most 2^24 voxels, in at most 1024 slabs
            assert vctx.variant == 4 and e.status == pcs.PCS_ERR_UNSUPPORTED and span >= 30000, (n, span, leaf)
            continue
        assert np.array_equal(got, R.voxel_merge(rec, leaf)), (n, span, leaf)
        exact += 1
    assert exact >= len(cases) - 2
    assert len(vctx.voxel_merge(np.zeros((0, 5), np.int16), 10)) == 0


def test_voxel_merge_skewed_and_unaligned(vctx, R):
    """Holes pile 1/16 of every camera's points into one voxel (SPEC.md s1): voxels that span many
    warps, chunks and sort tiles; plus a record pointer that is only 2-byte aligned."""
    rng = np.random.default_rng(13)
    n = 300000
    rec = random_records(rng, n)
    rec[:, :3] = rng.integers(-3000, 3000, (n, 3))
    rec[rng.random(n) < 0.6, :3] = (123, -457, 2000)          # one giant voxel
    rec[rng.random(n) < 0.1, :3] = (-32768, -32768, -32768)    # and one at the corner of the range
    want = R.voxel_merge(rec, 10)
    assert np.array_equal(vctx.voxel_merge(rec, 10), want)
    buf = torch.zeros(n * 5 + 8, dtype=torch.int16, device="cuda")
    out = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
    buf[1:1 + n * 5] = torch.from_numpy(rec.reshape(-1)).cuda()
    cs = torch.cuda.current_stream().cuda_stream
    nv = vctx.voxel_merge_dev(buf.data_ptr() + 2, n, 10, out.data_ptr(), cs)
    torch.cuda.synchronize()
    assert nv == len(want) and np.array_equal(out[: nv * 5].cpu().numpy().reshape(-1, 5), want)


@pytest.mark.parametrize("n_slabs", [1, 3, 8])
def test_voxel_slabs_concatenate_to_the_full_merge(ctx, R, n_slabs):
    """Sharded merge: equal-population z-slabs, merged independently, concatenated in slab order,
    are bit for bit the full merge (the multi-GPU path runs one slab per rank)."""
    rng = np.random.default_rng(15)
    for n, leaf, skew in [(200000, 10, False), (150001, 7, True), (5000, 1, False), (3, 10, False)]:
        rec = random_records(rng, n)
        rec[:, :3] = (rng.normal(0, 800, (n, 3))).clip(-32768, 32767).astype(np.int16)
        if skew:
            rec[rng.random(n) < 0.5, :3] = (5, 5, -1234)      # one voxel holds half the cloud
        want = R.voxel_merge(rec, leaf)
        d = torch.from_numpy(rec.reshape(-1)).cuda()
        out = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
        cs = torch.cuda.current_stream().cuda_stream
        splits, pts = ctx.voxel_slab_plan_dev(d.data_ptr(), n, leaf, n_slabs, cs)
        assert len(splits) == n_slabs + 1 and sum(pts) == n and all(a <= b for a, b in zip(splits, splits[1:]))
        kz = np.floor_divide(rec[:, 2].astype(np.int32), leaf)
        assert splits[0] == kz.min() and splits[-1] == kz.max() + 1
        got = []
        for r in range(n_slabs):
            assert pts[r] == int(((kz >= splits[r]) & (kz < splits[r + 1])).sum())
            nv = ctx.voxel_merge_slab_dev(d.data_ptr(), n, leaf, splits[r], splits[r + 1], out.data_ptr(), cs)
            torch.cuda.synchronize()
            got.append(out[: nv * 5].cpu().numpy().reshape(-1, 5).copy())
        assert np.array_equal(np.concatenate(got), want), (n, leaf, n_slabs)
        if not skew and n > 1000:
            assert max(pts) < 2.0 * n / n_slabs + 0.02 * n     # cuts are on whole planes, so only roughly equal
    # a slab with nothing in it
    assert ctx.voxel_merge_slab_dev(d.data_ptr(), n, 10, 3000, 3100, out.data_ptr(), cs) == 0


def test_voxel_merge_wide_keys_fall_back(vctx, R):
    """leaf 1 mm: 3 x 16 key bits + 18 index bits do not fit the one-sweep sort word."""
    rng = np.random.default_rng(14)
    n = 150000
    rec = random_records(rng, n)
    if vctx.variant in (2, 3, 4):
        with pytest.raises(pcs.PcsError) as e:
            vctx.voxel_merge(rec, 1)
        assert e.value.status == pcs.PCS_ERR_UNSUPPORTED
    else:
        assert np.array_equal(vctx.voxel_merge(rec, 1), R.voxel_merge(rec, 1))


def test_voxel_merge_full_size_properties(vctx):
    # 4 cameras x 1280x720 of plausible geometry: idempotence, sortedness, bounds
    rng = np.random.default_rng(12)
    n = 4 * 921600
    rec = random_records(rng, n)
    rec[:, :3] = (rng.normal(0, 1500, (n, 3))).clip(-32000, 32000).astype(np.int16)
    m = vctx.voxel_merge(rec, 10)
    k = np.floor_divide(m[:, :3].astype(np.int64), 10)
    key = (k[:, 2] << 40) + (k[:, 1] << 20) + k[:, 0]
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(vctx.voxel_merge(m, 10), m)
    assert len(m) == len(np.unique(np.floor_divide(rec[:, :3].astype(np.int32), 10), axis=0))


@pytest.mark.parametrize("n_cams,w,h", [(4, 1280, 720), (20, 848, 480)], ids=["config3_4x720p", "config5_20x480p"])
def test_baseline_configs_k1_concat_voxel(ctx, R, n_cams, w, h):
    """BASELINE.json configs #3 / #5 on one GPU: every camera's K1 writes straight into its slot of
    the stitched buffer (the concat costs nothing), then the 1 cm voxel merge runs on the stitched
    records.  Compared with oracle deproject -> pack -> concat -> voxel merge, bit for bit."""
    c2 = pcs.Context(device=0, max_streams=n_cams)
    total = n_cams * w * h
    st = torch.zeros(16 + total * 10, dtype=torch.uint8, device="cuda")
    hdr = np.frombuffer(np.int32(total * 10).tobytes(), np.uint8).copy()
    st[12:16] = torch.from_numpy(hdr).cuda()
    jobs, keep, want_pay = [], [], []
    cal = oracle.make_calib(w, h, translation=synth.D2C_BASELINE)
    for cam in range(n_cams):
        c2.set_stream(cam, pcs.stream_desc(w, h, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE))
        z, col = synth.depth_frame(w, h, cam, 0), synth.color_frame(w, h, cam, 0)
        dz, dc = dev(z), dev(col)
        keep.append((dz, dc))
        jobs.append((cam, dz.data_ptr(), dc.data_ptr(), st.data_ptr() + 16 + cam * w * h * 10))
        want_pay.append(R.frame(cal, z, col, 3, w * 3, synth.TF_STITCH[cam % 8]))
    cs = torch.cuda.current_stream().cuda_stream
    b = c2.batch(jobs)
    b.run(cs)
    torch.cuda.synchronize()
    want_stitched = R.concat(want_pay, 1)
    assert np.array_equal(st[12:].cpu().numpy(), want_stitched)
    out = torch.zeros(total * 5, dtype=torch.int16, device="cuda")
    nv = c2.voxel_merge_dev(st.data_ptr() + 16, total, 10, out.data_ptr(), cs)
    torch.cuda.synchronize()
    want_vox = R.voxel_merge(np.concatenate(want_pay), 10)
    assert nv == len(want_vox)
    assert np.array_equal(out[: nv * 5].cpu().numpy().reshape(-1, 5), want_vox)
    b.close()
    c2.close()


# ---- the whole path in one call: host frames in, the reference's stitched buffer out ------------------
@pytest.mark.parametrize("downsample", [1, 2, 3])
def test_stitch_frames_host_api_vs_oracle(R, downsample):
    """pcs_b200_stitch_frames == n x (calculate + sendXYZRGBPointcloud) + readCloud + sendStitchToUnity's concat
    (src/pcs-camera-optimized.cpp:288-292, src/pcs-multicamera-client.cpp:363-395), mixed geometries."""
    geoms = [(1280, 720, synth.D2C_BASELINE), (848, 480, (0.0, 0.0, 0.0)), (1280, 720, (0.0, 0.0, 0.0)),
             (640, 480, synth.D2C_BASELINE)]
    c = pcs.Context(device=0, max_streams=len(geoms))
    zs, cols, want = [], [], []
    for cam, (w, h, tr) in enumerate(geoms):
        c.set_stream(cam, pcs.stream_desc(w, h, tf=synth.TF_STITCH[cam], translation=tr))
        z, col = synth.depth_frame(w, h, cam, 3), synth.color_frame(w, h, cam, 3)
        zs.append(z)
        cols.append(col)
        want.append(R.frame(oracle.make_calib(w, h, translation=tr), z, col, 3, w * 3, synth.TF_STITCH[cam]).reshape(-1))
    order = [2, 0, 3, 1]     # stitched order is the order of `streams`, not of the stream ids
    got = c.stitch_frames(order, [zs[i] for i in order], [cols[i] for i in order], downsample)
    assert np.array_equal(got, R.concat([want[i] for i in order], downsample))
    # two frames in flight on two slots, and a slot reused for another camera set
    bufs = [np.zeros(4 + 10 * sum(z.size for z in zs), np.uint8) for _ in range(2)]
    c.stitch_frames_begin(0, [0, 1], zs[:2], cols[:2], bufs[0], downsample)
    c.stitch_frames_begin(1, [3], zs[3:], cols[3:], bufs[1], downsample)
    n0, n1 = c.stitch_frames_end(0), c.stitch_frames_end(1)
    assert np.array_equal(bufs[0][: n0 + 4], R.concat(want[:2], downsample))
    assert np.array_equal(bufs[1][: n1 + 4], R.concat(want[3:], downsample))
    with pytest.raises(pcs.PcsError):
        c.stitch_frames_end(0)                      # nothing in flight
    with pytest.raises(pcs.PcsError):
        c.stitch_frames([0], zs[:1], cols[:1], 1, stitched=np.zeros(100, np.uint8))   # too small
    c.close()


def test_voxel_merge_msd_shapes(R):
    """The MSD partition's corner cases: dense planes (thousands of points per z plane and per key row), a
    cloud that is one z plane, one voxel row, leaf sizes up to 32 mm, a sub-bucket that needs several
    accumulator rounds, and the z-slab filter with empty slabs inside the range."""
    c = pcs.Context(device=0, max_streams=1, voxel_variant=4)
    rng = np.random.default_rng(21)

    def check(rec, leaf):
        assert np.array_equal(c.voxel_merge(rec, leaf), R.voxel_merge(rec, leaf)), (len(rec), leaf)

    n = 400000
    # a wall in one z plane (x, y spread; ~ 1 point per voxel) + a floor (one y row band) + a dense blob
    rec = random_records(rng, n)
    rec[:, 0] = rng.integers(-2000, 2000, n)
    rec[:, 1] = rng.integers(-1200, 1200, n)
    rec[:, 2] = 1503
    rec[: n // 4, 1] = -1195
    rec[: n // 4, 2] = rng.integers(500, 4000, n // 4)
    rec[n // 4: n // 4 + 50000, :3] = rng.integers(100, 160, (50000, 3))      # 6^3 voxels, ~230 points each
    check(rec, 10)
    # one voxel row, every voxel hit ~100 times
    rec = random_records(rng, 100000)
    rec[:, 0] = rng.integers(-5000, 5000, 100000)
    rec[:, 1:3] = (77, -3)
    check(rec, 10)
    # leaf sizes: offsets of 1 .. 5 bits
    for leaf in (1, 2, 5, 16, 17, 32):
        rec = random_records(rng, 60000)
        rec[:, :3] = rng.integers(-40 * leaf, 40 * leaf, (60000, 3))
        check(rec, leaf)
    with pytest.raises(pcs.PcsError):
        c.voxel_merge(rec, 33)
    # many voxels in one sub-bucket range (several accumulator rounds) next to empty space
    rec = random_records(rng, 30000)
    rec[:, 0] = rng.integers(0, 1270, 30000)
    rec[:, 1] = rng.integers(0, 100, 30000)
    rec[:, 2] = 0
    rec[-1, :3] = (30000, 30000, 3000)        # stretches the box
    check(rec, 10)
    c.close()


def test_voxel_merge_beyond_the_uint32_sum_limit(R):
    """20 cameras x 1280x720 = 18.4 M points: 255 * n no longer fits the uint32 colour sums of the sort-based
    variants; auto then takes the slab-partition variant, which counts colours as two 4-bit halves where a voxel is
    that crowded."""
    c = pcs.Context(device=0, max_streams=1)
    rng = np.random.default_rng(31)
    n = 20 * 921600
    rec = random_records(rng, n)
    rec[:, :3] = (rng.normal(0, 1200, (n, 3))).clip(-32000, 32000).astype(np.int16)
    rec[: n // 2, :3] = (40, -7, 1999)            # 9.2 M points in one voxel: its colour sums pass 2^31
    d = torch.from_numpy(rec.reshape(-1)).cuda()
    out = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
    nv = c.voxel_merge_dev(d.data_ptr(), n, 10, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = R.voxel_merge(rec, 10)
    assert nv == len(want) and np.array_equal(out[: nv * 5].cpu().numpy().reshape(-1, 5), want)
    c.close()


def test_voxel_merge_async_queue_of_frames(R):
    """pcs_b200_voxel_merge_async_dev: several merges queued back to back on one stream, one synchronisation;
    counts and voxels per frame against the oracle; a cloud whose sort word does not fit reports
    PCS_ERR_UNSUPPORTED in its count."""
    c = pcs.Context(device=0, max_streams=1)
    rng = np.random.default_rng(41)
    cs = torch.cuda.current_stream().cuda_stream
    clouds, leaves = [], [10, 10, 7, 10, 1]
    for k, n in enumerate([120000, 5, 80001, 300000, 150000]):
        rec = random_records(rng, n)
        span = 32768 if k == 4 else 1500 + 500 * k            # the last one: full range at leaf 1 -> the word does not fit
        rec[:, :3] = rng.integers(-span, span, (n, 3))
        if k == 3:
            rec[rng.random(n) < 0.4, :3] = (9, 9, 9)
        clouds.append(rec)
    d_in = [torch.from_numpy(r.reshape(-1)).cuda() for r in clouds]
    d_out = [torch.zeros(r.size, dtype=torch.int16, device="cuda") for r in clouds]
    counts = torch.full((len(clouds),), -77, dtype=torch.int32, device="cuda")
    for k, rec in enumerate(clouds):
        c.voxel_merge_async_dev(d_in[k].data_ptr(), len(rec), leaves[k], d_out[k].data_ptr(), counts.data_ptr() + 4 * k, cs)
    torch.cuda.synchronize()
    got = counts.cpu().numpy()
    for k, rec in enumerate(clouds[:4]):
        want = R.voxel_merge(rec, leaves[k])
        assert got[k] == len(want), (k, got[k], len(want))
        assert np.array_equal(d_out[k][: got[k] * 5].cpu().numpy().reshape(-1, 5), want), k
    assert got[4] == pcs.PCS_ERR_UNSUPPORTED
    # a slab of cloud 0, asynchronously
    kz = np.floor_divide(clouds[0][:, 2].astype(np.int32), 10)
    lo, hi = int(np.percentile(kz, 30)), int(np.percentile(kz, 70))
    c.voxel_merge_slab_async_dev(d_in[0].data_ptr(), len(clouds[0]), 10, lo, hi, d_out[0].data_ptr(), counts.data_ptr(), cs)
    torch.cuda.synchronize()
    want = R.voxel_merge(clouds[0][(kz >= lo) & (kz < hi)], 10)
    nv = int(counts[0].item())
    assert nv == len(want) and np.array_equal(d_out[0][: nv * 5].cpu().numpy().reshape(-1, 5), want)
    c.close()


@pytest.mark.parametrize("n_ranks", [1, 3, 8])
def test_shard_before_exchange_kernels_on_one_gpu(R, n_ranks):
    """The all-to-all of the sharded multi-GPU merge (pcs_b200_shard_*) with every "rank" living on this GPU: each
    rank's records are histogrammed, every rank derives the cuts from all histograms, the scatter fills the ranks'
    inboxes, every inbox is merged with a device-side point count -- slabs in rank order == the merge of everything.
    (tests/test_multigpu_gpu.py runs the same through NVLink on two GPUs.)"""
    c = pcs.Context(device=0, max_streams=1)
    rng = np.random.default_rng(50 + n_ranks)
    leaf = 10
    sizes = [int(rng.integers(20000, 90000)) for _ in range(n_ranks)]
    if n_ranks > 2:
        sizes[1] = 0                                        # a rank without points
    recs = []
    for k, n in enumerate(sizes):
        r = random_records(rng, n)
        r[:, :3] = (rng.normal(0, 700, (n, 3)) + (300 * k, 0, 150 * k)).clip(-32000, 32000).astype(np.int16)
        if k == 0 and n:
            r[: n // 3, :3] = (11, 12, -513)                # a crowded voxel
        recs.append(r)
    total = sum(sizes)
    zbins = pcs.lib.pcs_b200_shard_zbins(leaf)
    cs = torch.cuda.current_stream().cuda_stream
    d_rec = [torch.from_numpy(r.reshape(-1).copy()).cuda() if len(r) else torch.zeros(8, dtype=torch.int16, device="cuda") for r in recs]
    zh = [torch.full((zbins,), 7, dtype=torch.int32, device="cuda") for _ in range(n_ranks)]
    cur = [torch.full((1,), 99, dtype=torch.int32, device="cuda") for _ in range(n_ranks)]
    inbox = [torch.zeros(total * 5 + 8, dtype=torch.int16, device="cuda") for _ in range(n_ranks)]
    peers = pcs.ShardPeers()
    peers.n_ranks, peers.rank, peers.capacity_records = n_ranks, 0, total
    for r in range(n_ranks):
        peers.inbox_dev[r], peers.cursor_dev[r], peers.zhist_dev[r] = inbox[r].data_ptr(), cur[r].data_ptr(), zh[r].data_ptr()
    for r in range(n_ranks):
        c.shard_hist_dev(d_rec[r].data_ptr(), sizes[r], leaf, zh[r].data_ptr(), cur[r].data_ptr(), cs)
    splits = torch.zeros(n_ranks + 1, dtype=torch.int32, device="cuda")
    zslab = torch.zeros((zbins + 15) & ~15, dtype=torch.uint8, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    c.shard_plan_dev(peers, leaf, splits.data_ptr(), zslab.data_ptr(), cs)
    for r in range(n_ranks):
        c.shard_scatter_dev(d_rec[r].data_ptr(), sizes[r], leaf, zslab.data_ptr(), peers, err.data_ptr(), cs)
    out = torch.zeros(total * 5, dtype=torch.int16, device="cuda")
    cnt = torch.zeros(n_ranks, dtype=torch.int32, device="cuda")
    got = []
    for r in range(n_ranks):
        c.voxel_merge_counted_async_dev(inbox[r].data_ptr(), total, cur[r].data_ptr(), leaf, out.data_ptr(), cnt.data_ptr() + 4 * r, cs)
        torch.cuda.synchronize()
        nv = int(cnt[r].item())
        assert nv >= 0
        got.append(out[: nv * 5].cpu().numpy().reshape(-1, 5).copy())
    assert int(err.item()) == 0
    allrec = np.concatenate(recs)
    kz = np.floor_divide(allrec[:, 2].astype(np.int32), leaf)
    sp = splits.cpu().numpy()
    assert sp[0] == kz.min() and sp[-1] == kz.max() + 1 and np.all(np.diff(sp) >= 0)
    fill = [int(x.item()) for x in cur]
    assert fill == [int(((kz >= sp[r]) & (kz < sp[r + 1])).sum()) for r in range(n_ranks)] and sum(fill) == total
    if n_ranks > 1:
        assert max(fill) < 2.0 * total / n_ranks + 0.4 * total      # equal population up to whole planes (one is crowded)
    assert np.array_equal(np.concatenate(got), R.voxel_merge(allrec, leaf))
    c.close()


@pytest.mark.parametrize("variant", [0, 4], ids=["sort_with_merged_runs", "msd"])
def test_voxel_merge_of_a_smooth_scene(R, variant):
    """A surface as a depth camera sees it (5 mm noise): raster neighbours share voxels, so the sort merges the runs
    inside every thread's eight records before sorting (sw_plan decides on the device).  K1 records of two cameras,
    the full cloud, a z slab of it, and the asynchronous call -- all against the oracle."""
    w, h, cams = 848, 480, 2
    c = pcs.Context(device=0, max_streams=cams, voxel_variant=variant)
    cal = oracle.make_calib(w, h, translation=synth.D2C_BASELINE)
    recs = [R.frame(cal, synth.depth_frame(w, h, cam, 1, noise_mm=synth.SMOOTH_NOISE_MM), synth.color_frame(w, h, cam, 1), 3,
                    w * 3, synth.TF_STITCH[cam]) for cam in range(cams)]
    rec = np.concatenate(recs)
    want = R.voxel_merge(rec, 10)
    assert len(want) < 0.6 * len(rec)                     # several points per voxel
    n = len(rec)
    d = torch.from_numpy(rec.reshape(-1).copy()).cuda()
    out = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
    cs = torch.cuda.current_stream().cuda_stream
    nv = c.voxel_merge_dev(d.data_ptr(), n, 10, out.data_ptr(), cs)
    torch.cuda.synchronize()
    assert nv == len(want) and np.array_equal(out[: nv * 5].cpu().numpy().reshape(-1, 5), want)
    kz = np.floor_divide(rec[:, 2].astype(np.int32), 10)
    lo, hi = int(np.percentile(kz, 25)), int(np.percentile(kz, 80))
    nv = c.voxel_merge_slab_dev(d.data_ptr(), n, 10, lo, hi, out.data_ptr(), cs)
    torch.cuda.synchronize()
    want_slab = R.voxel_merge(rec[(kz >= lo) & (kz < hi)], 10)
    assert nv == len(want_slab) and np.array_equal(out[: nv * 5].cpu().numpy().reshape(-1, 5), want_slab)
    if variant == 0:
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        c.voxel_merge_async_dev(d.data_ptr(), n, 10, out.data_ptr(), cnt.data_ptr(), cs)
        torch.cuda.synchronize()
        assert int(cnt.item()) == len(want) and np.array_equal(out[: len(want) * 5].cpu().numpy().reshape(-1, 5), want)
        # other leaf sizes with merged runs: 3 mm (2-bit offsets) and 40 mm (6-bit offsets: the widest gather word)
        for leaf in (3, 40):
            got = c.voxel_merge(rec, leaf)
            assert np.array_equal(got, R.voxel_merge(rec, leaf)), leaf
    c.close()


@pytest.mark.parametrize("downsample", [1, 3])
def test_stitch_frames_with_cutoff_streams(R, downsample):
    """-c on the camera (src/pcs-camera-optimized.cpp:499-577) + the stitcher's concat: the compacted record counts only exist
    on the device, the concat reads them there.  One -c camera with the reference's reversed mask lanes, one with straight
    lanes, one without -c."""
    w, h = 848, 480
    c = pcs.Context(device=0, max_streams=3)
    cal = oracle.make_calib(w, h, translation=synth.D2C_BASELINE)
    zs, cols, want = [], [], []
    for cam, (cut, rev) in enumerate([(True, True), (False, True), (True, False)]):
        d = pcs.stream_desc(w, h, tf=synth.TF_CAMERA, translation=synth.D2C_BASELINE, cutoff=cut, lane_reversed=rev)
        c.set_stream(cam, d)
        z, col = synth.depth_frame(w, h, cam, 2, lo=300, hi=2600), synth.color_frame(w, h, cam, 2)
        zs.append(z)
        cols.append(col)
        xyz, uv = R.deproject(cal, z)
        if cut and rev:
            rec = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA, cutoff=True)
        else:
            rec = R.pack(xyz, uv, col, w, h, 3, w * 3, synth.TF_CAMERA)
            if cut:
                rec = rec[(xyz[:, 2] > 0) & (xyz[:, 2] <= 1.5) & (xyz[:, 0] > -2) & (xyz[:, 0] <= 2)]
        want.append(rec.reshape(-1))
    assert 1000 < want[0].size // 5 < w * h and want[1].size // 5 == w * h
    for rep in range(2):                      # the slot is reused
        got = c.stitch_frames([0, 1, 2], zs, cols, downsample)
        assert np.array_equal(got, R.concat(want, downsample))
    c.close()
