"""N > 1 host logic on CPU: world_size-2 gloo processes shard the cameras, fill their slots of
the replicated stitched buffer (records come from the oracle here -- there is no GPU), run the
in-place exchange and must end up with the reference's stitched layout
(src/pcs-multicamera-client.cpp:385-395) byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointcloud_stitching_b200 import multigpu, synth


def test_partition_and_layout():
    assert multigpu.partition(20, 8) == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11], [12, 13], [14, 15],
                                         [16, 17], [18, 19]]
    assert multigpu.partition(8, 8) == [[k] for k in range(8)]
    assert multigpu.partition(3, 4) == [[0], [1], [2], []]
    L = multigpu.StitchLayout([407040] * 20, 8)
    assert L.total_bytes == 81408000 and not L.equal
    assert L.rank_offset[4] == 12 * 4070400 and L.rank_bytes[4] == 2 * 4070400
    assert L.rank_of(13) == 4
    assert multigpu.StitchLayout([921600] * 8, 8).equal
    with pytest.raises(ValueError):
        multigpu.StitchLayout([921600] * 300, 8)      # > int32 header


def test_partition_and_layout_properties():
    """For any camera count / world size: contiguous blocks in camera order, sizes differ by at most one,
    rank byte ranges tile the payload exactly, and every camera has exactly one owner."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 64), st.integers(1, 16), st.integers(1, 4096))
    def check(n_cams, world, pts):
        blocks = multigpu.partition(n_cams, world)
        assert len(blocks) == world and [c for b in blocks for c in b] == list(range(n_cams))
        sizes = [len(b) for b in blocks]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
        L = multigpu.StitchLayout([pts * 8] * n_cams, world)
        assert L.total_bytes == n_cams * pts * 80
        pos = 0
        for r in range(world):
            if L.rank_bytes[r]:
                assert L.rank_offset[r] == pos
            pos += L.rank_bytes[r]
        assert pos == L.total_bytes
        assert all(L.rank_of(c) == r for r, b in enumerate(blocks) for c in b)
        assert L.equal == (n_cams % world == 0)

    check()


def test_pull_exchange_job_table():
    """The pull exchange's host logic: frame slots do not overlap, and every rank's job table names
    every camera once per frame, reading the owner's allocation and writing the local stitched slot."""
    W, H, F = 64, 8, 3
    L = multigpu.StitchLayout([W * H] * 5, 2)                  # ranks own cams [0,1,2] and [3,4]
    depth_off, color_off, total = multigpu.frame_slots(W, H, W * 3, 3, F)
    spans = []
    for lc in range(3):
        for f in range(F):
            spans += [(depth_off(lc, f), W * H * 2), (color_off(lc, f), H * W * 3)]
    spans.sort()
    assert all(o % 256 == 0 for o, _ in spans)
    assert all(a + n <= b for (a, n), (b, _) in zip(spans, spans[1:])) and spans[-1][0] + spans[-1][1] <= total
    for rank in range(2):
        fs = multigpu.SymmetricFrameSet.__new__(multigpu.SymmetricFrameSet)
        fs.layout, fs.rank, fs.n_frames = L, rank, F
        fs.depth_off, fs.color_off, fs.nbytes = depth_off, color_off, total
        fs.bases = [0x10000000, 0x20000000]
        stitched = [multigpu.StitchedBuffer(L, rank, "cpu") for _ in range(F)]
        jobs = fs.pull_jobs(stitched)
        assert [j[0] for j in jobs] == list(range(5)) * F
        for k, (cam, z, c, out, _, _, flags) in enumerate(jobs):
            f, owner = k // 5, L.rank_of(cam)
            lc = L.cams_of[owner].index(cam)
            assert z == fs.bases[owner] + depth_off(lc, f) and c == fs.bases[owner] + color_off(lc, f)
            assert out == stitched[f].slot_ptr(cam)
            assert flags == (0 if owner == rank else 1)      # PCS_B200_JOB_REMOTE_FRAME on the peers' frames


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, points, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    R = oracle.restatement()
    layout = multigpu.StitchLayout(points, world)
    buf = multigpu.StitchedBuffer(layout, rank, "cpu")
    for cam in buf.my_cams():
        w, h = points[cam] // 8, 8
        rec = R.frame(oracle.make_calib(w, h, translation=synth.D2C_BASELINE), synth.depth_frame(w, h, cam, 0),
                      synth.color_frame(w, h, cam, 0), 3, w * 3, synth.TF_STITCH[cam % 8])
        buf.slot(cam).copy_(torch.from_numpy(rec.reshape(-1).view(np.uint8).copy()))
    buf.gather()
    np.save(os.path.join(result_dir, "rank%d.npy" % rank), buf.wire_bytes().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("points", [[512, 512, 512, 512], [512, 256, 1024, 64, 128]], ids=["equal", "ragged"])
def test_two_rank_gloo_exchange_matches_reference_layout(tmp_path, points, restatement):
    import oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), points, str(tmp_path)), nprocs=world, join=True)
    payloads = []
    for cam, n in enumerate(points):
        w, h = n // 8, 8
        payloads.append(restatement.frame(oracle.make_calib(w, h, translation=synth.D2C_BASELINE),
                                          synth.depth_frame(w, h, cam, 0), synth.color_frame(w, h, cam, 0), 3, w * 3,
                                          synth.TF_STITCH[cam % 8]))
    want = restatement.concat(payloads, 1)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert np.array_equal(got, want), "rank %d" % r
