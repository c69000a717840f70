"""-m gpu, needs >= 2 GPUs (gpurun --gpus 2): one process per GPU over NCCL.  All exchange paths
-- K1 pulling the peers' raw frames over NVLink (SymmetricFrameSet), the fused K1 + peer-store
kernel (pcs_b200_batch_create_fanout) and the NCCL all-gather baseline -- must leave the
reference's stitched layout, bit-exact against the oracle, on every rank."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices", allow_module_level=True)

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

W, H = 256, 48


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cams_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pointcloud_stitching_b200 as pcs
    from pointcloud_stitching_b200 import multigpu, synth
    dev = torch.device("cuda", rank)
    layout = multigpu.StitchLayout([W * H] * cams_total, world)
    ctx = pcs.Context(device=rank, max_streams=cams_total)
    n_frames = 2
    sset = multigpu.SymmetricStitchedSet(layout, rank, dev, n_frames)
    plain = [multigpu.StitchedBuffer(layout, rank, dev) for _ in range(n_frames)]
    keep, jobs_fused, jobs_plain = [], [], []
    for cam in layout.cams_of[rank]:
        ctx.set_stream(cam, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE))
        for f in range(n_frames):
            z = torch.from_numpy(synth.depth_frame(W, H, cam, f).view(np.int16)).to(dev)
            c = torch.from_numpy(synth.color_frame(W, H, cam, f)).to(dev)
            keep.append((z, c))
            jobs_fused.append((cam, z.data_ptr(), c.data_ptr(), sset.frames[f].slot_ptr(cam)))
            jobs_plain.append((cam, z.data_ptr(), c.data_ptr(), plain[f].slot_ptr(cam)))
    cs = torch.cuda.current_stream().cuda_stream
    # fused: one kernel computes and stores to every rank's mirror
    bf = ctx.batch_fanout(jobs_fused, sset.local_base, sset.nbytes, sset.peer_bases)
    sset.barrier()
    bf.run(cs)
    sset.barrier()
    torch.cuda.synchronize()
    # baseline: K1 then NCCL
    bp = ctx.batch(jobs_plain)
    bp.run(cs)
    for f in range(n_frames):
        plain[f].gather()
    torch.cuda.synchronize()
    # pull: every rank computes every camera, the peers' frames come through NVLink peer memory
    fset = multigpu.SymmetricFrameSet(layout, rank, dev, W, H, n_frames)
    pulled = [multigpu.StitchedBuffer(layout, rank, dev) for _ in range(n_frames)]
    for cam in range(cams_total):
        ctx.set_stream(cam, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE))
    for cam in layout.cams_of[rank]:
        for f in range(n_frames):
            fset.upload(cam, f, synth.depth_frame(W, H, cam, f), synth.color_frame(W, H, cam, f))
    torch.cuda.synchronize()
    fset.barrier()
    bq = ctx.batch(fset.pull_jobs(pulled))
    assert bq.launches == 1
    bq.run(cs)
    fset.barrier()
    torch.cuda.synchronize()
    # sharded voxel merge of the pulled cloud: one z-slab per rank, then exchanged
    n_pts = layout.total_bytes // 10
    vox = torch.zeros(n_pts * 5, dtype=torch.int16, device=dev)
    nv, mine = multigpu.sharded_voxel_merge(ctx, pulled[0].payload.data_ptr(), n_pts, 10, rank, world, vox, cs)
    torch.cuda.synchronize()
    assert 0 < mine < nv
    np.save(os.path.join(out_dir, "vox_r%d.npy" % rank), vox[: nv * 5].cpu().numpy().reshape(-1, 5))
    for f in range(n_frames):
        np.save(os.path.join(out_dir, "pull_r%d_f%d.npy" % (rank, f)), pulled[f].wire_bytes().cpu().numpy())
        np.save(os.path.join(out_dir, "fused_r%d_f%d.npy" % (rank, f)), sset.frames[f].wire_bytes().cpu().numpy())
        np.save(os.path.join(out_dir, "nccl_r%d_f%d.npy" % (rank, f)), plain[f].wire_bytes().cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cams_total", [4, 5], ids=["equal", "ragged"])
def test_pull_fused_and_nccl_exchange_match_oracle(tmp_path, cams_total):
    import oracle
    from pointcloud_stitching_b200 import synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), cams_total, str(tmp_path)), nprocs=world, join=True)
    R = oracle.restatement()
    cal = oracle.make_calib(W, H, translation=synth.D2C_BASELINE)
    for f in range(2):
        want = R.concat([R.frame(cal, synth.depth_frame(W, H, cam, f), synth.color_frame(W, H, cam, f), 3, W * 3,
                                 synth.TF_STITCH[cam % 8]) for cam in range(cams_total)], 1)
        for r in range(world):
            for kind in ("pull", "fused", "nccl"):
                got = np.load(os.path.join(str(tmp_path), "%s_r%d_f%d.npy" % (kind, r, f)))
                assert np.array_equal(got, want), (kind, r, f)
        if f == 0:
            want_vox = R.voxel_merge(want[4:].view(np.int16).reshape(-1, 5), 10)
            for r in range(world):
                assert np.array_equal(np.load(os.path.join(str(tmp_path), "vox_r%d.npy" % r)), want_vox), r


def test_single_process_pull_over_peer_access():
    """A C++-style host: one process, one context per GPU, plain cudaMalloc frames.  After
    pcs_b200_enable_peer every context deprojects ALL cameras -- its own frames from HBM, the other
    GPU's over NVLink -- and both stitched buffers must equal the oracle's."""
    import oracle
    import pointcloud_stitching_b200 as pcs
    from pointcloud_stitching_b200 import multigpu, synth
    R = oracle.restatement()
    cams, world = 4, 2
    layout = multigpu.StitchLayout([W * H] * cams, world)
    cal = oracle.make_calib(W, H, translation=synth.D2C_BASELINE)
    want = R.concat([R.frame(cal, synth.depth_frame(W, H, cam, 0), synth.color_frame(W, H, cam, 0), 3, W * 3,
                             synth.TF_STITCH[cam % 8]) for cam in range(cams)], 1)
    frames = {}
    for r in range(world):
        for cam in layout.cams_of[r]:
            frames[cam] = (torch.from_numpy(synth.depth_frame(W, H, cam, 0).view(np.int16)).to("cuda:%d" % r),
                           torch.from_numpy(synth.color_frame(W, H, cam, 0)).to("cuda:%d" % r))
    for r in range(world):
        torch.cuda.synchronize(r)
    ctxs, bufs = [], []
    for r in range(world):
        ctx = pcs.Context(device=r, max_streams=cams)
        ctx.enable_peer(1 - r)
        ctx.enable_peer(1 - r)          # idempotent
        for cam in range(cams):
            ctx.set_stream(cam, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE))
        buf = multigpu.StitchedBuffer(layout, r, torch.device("cuda", r))
        batch = ctx.batch([(cam, frames[cam][0].data_ptr(), frames[cam][1].data_ptr(), buf.slot_ptr(cam))
                           for cam in range(cams)])
        with torch.cuda.device(r):
            batch.run(torch.cuda.current_stream(r).cuda_stream)
        ctxs.append((ctx, batch))
        bufs.append(buf)
    for r in range(world):
        torch.cuda.synchronize(r)
        assert np.array_equal(bufs[r].wire_bytes().cpu().numpy(), want), r


def _shard_worker(rank, world, port, cams_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pointcloud_stitching_b200 as pcs
    from pointcloud_stitching_b200 import multigpu, synth
    dev = torch.device("cuda", rank)
    layout = multigpu.StitchLayout([W * H] * cams_total, world)
    mine = layout.cams_of[rank]
    ctx = pcs.Context(device=rank, max_streams=cams_total)
    n_frames, n_total = 3, cams_total * W * H
    cs = torch.cuda.current_stream().cuda_stream
    own = [torch.zeros(len(mine) * W * H * 5, dtype=torch.int16, device=dev) for _ in range(n_frames)]
    keep, jobs = [], []
    for i, cam in enumerate(mine):
        ctx.set_stream(cam, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE))
        for f in range(n_frames):
            z = torch.from_numpy(synth.depth_frame(W, H, cam, f).view(np.int16)).to(dev)
            c = torch.from_numpy(synth.color_frame(W, H, cam, f)).to(dev)
            keep.append((z, c))
            jobs.append((cam, z.data_ptr(), c.data_ptr(), own[f].data_ptr() + i * W * H * 10))
    batch = ctx.batch(jobs)
    sm = multigpu.ShardedMerge(ctx, rank, world, dev, n_total, 10, n_slots=2)
    outs = [torch.zeros(n_total * 5, dtype=torch.int16, device=dev) for _ in range(n_frames)]
    for rep in range(2):                 # twice: the slots (inbox, cursor, histogram) are reused
        batch.run(cs)
        for f in range(n_frames):
            sm.run(f % 2, own[f].data_ptr(), len(mine) * W * H, outs[f], cs)
            if f % 2 == 1 or f == n_frames - 1:
                torch.cuda.synchronize()             # slot results are read before the slot is reused
                for g in range(f - (f % 2), f + 1):
                    nv = int(sm.count[g % 2].item())
                    assert nv >= 0 and int(sm.err[g % 2].item()) == 0
                    np.save(os.path.join(out_dir, "slab_r%d_f%d.npy" % (rank, g)), outs[g][: nv * 5].cpu().numpy().reshape(-1, 5))
                    np.save(os.path.join(out_dir, "cuts_r%d_f%d.npy" % (rank, g)), sm.splits[g % 2].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cams_total", [4, 5], ids=["equal", "ragged"])
def test_sharded_merge_all_to_all_matches_the_single_gpu_merge(tmp_path, cams_total):
    """Shard-before-exchange: no rank ever holds the stitched cloud; slab r of the merged grid ends on rank r and the
    slabs in rank order are the oracle's merge of ALL cameras, bit for bit; every rank derives the same cuts."""
    import oracle
    from pointcloud_stitching_b200 import synth
    world = 2
    mp.spawn(_shard_worker, args=(world, _free_port(), cams_total, str(tmp_path)), nprocs=world, join=True)
    R = oracle.restatement()
    cal = oracle.make_calib(W, H, translation=synth.D2C_BASELINE)
    for f in range(3):
        rec = np.concatenate([R.frame(cal, synth.depth_frame(W, H, cam, f), synth.color_frame(W, H, cam, f), 3, W * 3,
                                      synth.TF_STITCH[cam % 8]) for cam in range(cams_total)])
        want = R.voxel_merge(rec, 10)
        slabs = [np.load(os.path.join(str(tmp_path), "slab_r%d_f%d.npy" % (r, f))) for r in range(world)]
        cuts = [np.load(os.path.join(str(tmp_path), "cuts_r%d_f%d.npy" % (r, f))) for r in range(world)]
        assert np.array_equal(cuts[0], cuts[1])
        assert np.array_equal(np.concatenate(slabs), want), f
        kz = np.floor_divide(rec[:, 2].astype(np.int32), 10)
        assert cuts[0][0] == kz.min() and cuts[0][-1] == kz.max() + 1
        pts = [int(((kz >= cuts[0][r]) & (kz < cuts[0][r + 1])).sum()) for r in range(world)]
        assert max(pts) < 0.6 * len(rec)              # equal-population cuts (whole planes)
        assert all(len(s) > 0 for s in slabs)
