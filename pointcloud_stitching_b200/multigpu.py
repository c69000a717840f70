"""Multi-GPU stitch: one process per GPU, cameras sharded in contiguous blocks, every rank ends a
step holding the reference's stitched buffer of ALL cameras.

The reference fans N camera streams into one host over TCP and concatenates them in
camera-index order behind an int32 byte count
(src/pcs-multicamera-client.cpp:381-395).  Here K1 writes every camera's records straight into
its slot of a replicated stitched buffer, so the "concat" is free, and the exchange is one of:

* pull (the product path, :class:`SymmetricFrameSet`): the raw z16 + RGB8 frames live in symmetric
  memory and every rank runs K1 over ALL cameras -- its own frames from HBM, the peers' through the
  kernel's TMA loads from NVLink peer memory.  5 B/pt cross the link instead of 10 B/pt, the transfer
  is the compute kernel's own input pipeline, and nobody writes into another rank's memory;
* fused push (:class:`SymmetricStitchedSet`): the stitched buffers live in symmetric memory and
  every rank's K1 stores each tile of records to its own copy AND to the same offset of every peer's
  copy (TMA bulk stores over NVLink, ``pcs_b200_batch_create_fanout``);
* NCCL baseline, equal shards  -> ``all_gather_into_tensor`` in place (send = recv + rank * count);
  ragged shards -> one ``broadcast`` per rank of that rank's slot (all-gather-v),
  e.g. 20 cameras on 8 GPUs = 3,3,3,3,2,2,2,2.

:func:`sharded_voxel_merge` then downsamples the stitched cloud with every rank merging one z-slab
of the voxel grid.  torch.distributed is plumbing only (rendezvous, symmetric-memory allocation,
NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RECORD_BYTES = 10
#: records start here so that they are 16-byte aligned; the int32 header sits at PAD - 4
PAD = 16


def partition(n_cams: int, world: int):
    """Contiguous blocks in camera order: the first n_cams % world ranks take one more."""
    base, extra = divmod(n_cams, world)
    out, start = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append(list(range(start, start + k)))
        start += k
    return out


class StitchLayout:
    """Byte offsets of every camera's slot in the stitched payload (camera-index order)."""

    def __init__(self, points_per_cam, world):
        self.points = [int(p) for p in points_per_cam]
        self.world = world
        self.cams_of = partition(len(self.points), world)
        self.cam_offset = np.concatenate([[0], np.cumsum([p * RECORD_BYTES for p in self.points])]).astype(np.int64)
        self.total_bytes = int(self.cam_offset[-1])
        self.rank_offset, self.rank_bytes = [], []
        for cams in self.cams_of:
            lo = int(self.cam_offset[cams[0]]) if cams else self.total_bytes
            hi = int(self.cam_offset[cams[-1] + 1]) if cams else self.total_bytes
            self.rank_offset.append(lo)
            self.rank_bytes.append(hi - lo)
        self.equal = len(set(self.rank_bytes)) == 1 and self.rank_bytes[0] > 0
        if self.total_bytes > 0x7FFFFFFF:
            raise ValueError("stitched payload exceeds the reference's int32 size header")

    def rank_of(self, cam):
        for r, cams in enumerate(self.cams_of):
            if cam in cams:
                return r
        raise IndexError(cam)


class StitchedBuffer:
    """A replicated stitched buffer ``[int32 bytes][cam0 records][cam1 records]...`` and its exchange."""

    def __init__(self, layout: StitchLayout, rank: int, device):
        self.layout, self.rank = layout, rank
        self.raw = torch.zeros(PAD + layout.total_bytes, dtype=torch.uint8, device=device)
        self.payload = self.raw[PAD:]
        hdr = np.frombuffer(np.int32(layout.total_bytes).tobytes(), np.uint8).copy()
        self.raw[PAD - 4:PAD] = torch.from_numpy(hdr).to(device)

    def slot(self, cam):
        """uint8 view of one camera's records."""
        o = self.layout.cam_offset
        return self.payload[int(o[cam]):int(o[cam + 1])]

    def slot_ptr(self, cam):
        return self.payload.data_ptr() + int(self.layout.cam_offset[cam])

    def my_cams(self):
        return self.layout.cams_of[self.rank]

    def gather(self, group=None):
        """In place: afterwards every rank holds every camera's records."""
        L = self.layout
        if L.world == 1:
            return
        if L.equal:
            n = L.rank_bytes[0]
            mine = self.payload[self.rank * n:(self.rank + 1) * n]
            dist.all_gather_into_tensor(self.payload, mine, group=group)
        else:
            for r in range(L.world):
                if L.rank_bytes[r]:
                    # `src` is a GLOBAL rank: translate the group-local index for groups other than WORLD
                    src = dist.get_global_rank(group, r) if group is not None else r
                    dist.broadcast(self.payload[L.rank_offset[r]:L.rank_offset[r] + L.rank_bytes[r]], src=src,
                                   group=group)

    def wire_bytes(self):
        """``[int32 bytes][records]`` exactly as the reference writes it to the viewer socket."""
        return self.raw[PAD - 4:]


class SymmetricStitchedSet:
    """``n_frames`` replicated stitched buffers in one symmetric-memory allocation, for the fused
    K1 + exchange kernel.  ``local_base`` / ``nbytes`` / ``peer_bases`` are what
    ``Context.batch_fanout`` takes; ``frames[f]`` are :class:`StitchedBuffer` views."""

    def __init__(self, layout: StitchLayout, rank: int, device, n_frames: int, group=None):
        import torch.distributed._symmetric_memory as symm
        self.layout, self.rank, self.n_frames = layout, rank, n_frames
        self.frame_bytes = (PAD + layout.total_bytes + 255) & ~255
        self.nbytes = self.frame_bytes * n_frames
        self.raw = symm.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.raw.zero_()
        self.handle = symm.rendezvous(self.raw, group if group is not None else dist.group.WORLD)
        ptrs = list(self.handle.buffer_ptrs)
        self.local_base = self.raw.data_ptr()
        assert ptrs[rank] == self.local_base
        self.peer_bases = [p for r, p in enumerate(ptrs) if r != rank]
        self.frames = []
        for f in range(n_frames):
            b = StitchedBuffer.__new__(StitchedBuffer)
            b.layout, b.rank = layout, rank
            b.raw = self.raw[f * self.frame_bytes:f * self.frame_bytes + PAD + layout.total_bytes]
            b.payload = b.raw[PAD:]
            hdr = np.frombuffer(np.int32(layout.total_bytes).tobytes(), np.uint8).copy()
            b.raw[PAD - 4:PAD] = torch.from_numpy(hdr).to(device)
            self.frames.append(b)

    def barrier(self):
        """Device-side barrier on the current stream: all ranks' fused launches have finished."""
        self.handle.barrier()


def frame_slots(width, height, stride, cams_per_rank_max, n_frames, color_height=None):
    """Byte offsets inside one rank's raw-frame allocation (the same on every rank):
    ``depth_off(local_cam, frame)``, ``color_off(local_cam, frame)`` and the total size.  Every frame is
    256-byte aligned (the pipelined kernel needs 16)."""
    dz = (width * height * 2 + 255) & ~255
    dc = ((height if color_height is None else color_height) * stride + 255) & ~255
    per_frame = dz + dc
    total = max(1, cams_per_rank_max) * n_frames * per_frame

    def depth_off(local_cam, frame):
        return (local_cam * n_frames + frame) * per_frame

    def color_off(local_cam, frame):
        return depth_off(local_cam, frame) + dz

    return depth_off, color_off, total


class SymmetricFrameSet:
    """The raw z16 depth + RGB8 colour frames of this rank's cameras, in symmetric memory, so that a
    peer's K1 can read them over NVLink.

    This is the *pull* exchange: instead of all-gathering the 10-byte records K1 produced, every
    rank runs K1 over ALL cameras -- its own from HBM, the peers' through TMA bulk loads from peer
    memory -- and writes the complete stitched buffer locally.  5 B/pt cross the link instead of
    10 B/pt, the transfer is the kernel's own input pipeline (all-gather fused into the compute
    kernel, tile by tile), and no rank ever writes to another rank's memory.  K1 is ~5x faster
    than the link, so recomputing a peer's records is cheaper than receiving them.
    """

    def __init__(self, layout: StitchLayout, rank: int, device, width, height, n_frames, stride=None, group=None,
                 color_height=None):
        import torch.distributed._symmetric_memory as symm
        self.layout, self.rank, self.n_frames = layout, rank, n_frames
        self.width, self.height = width, height
        self.stride = stride if stride is not None else width * 3
        cmax = max(len(c) for c in layout.cams_of)
        self.depth_off, self.color_off, self.nbytes = frame_slots(width, height, self.stride, cmax, n_frames,
                                                                  color_height)
        self.raw = symm.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.raw.zero_()
        self.handle = symm.rendezvous(self.raw, group if group is not None else dist.group.WORLD)
        self.bases = list(self.handle.buffer_ptrs)
        assert self.bases[rank] == self.raw.data_ptr()

    def upload(self, cam, frame, z16: np.ndarray, color: np.ndarray):
        """Place one of this rank's frames (host arrays) in its slot."""
        lc = self.layout.cams_of[self.rank].index(cam)
        dz = torch.from_numpy(np.ascontiguousarray(z16).view(np.uint8).reshape(-1))
        dc = torch.from_numpy(np.ascontiguousarray(color).reshape(-1))
        o = self.depth_off(lc, frame)
        self.raw[o:o + dz.numel()].copy_(dz)
        o = self.color_off(lc, frame)
        self.raw[o:o + dc.numel()].copy_(dc)

    def ptrs(self, cam, frame):
        """(z16_ptr, colour_ptr) of any camera's frame as seen from this rank: local HBM for its own
        cameras, NVLink peer memory otherwise."""
        owner = self.layout.rank_of(cam)
        lc = self.layout.cams_of[owner].index(cam)
        base = self.bases[owner]
        return base + self.depth_off(lc, frame), base + self.color_off(lc, frame)

    def pull_jobs(self, stitched_frames):
        """K1 jobs for EVERY camera and frame: (stream = camera index, z16, colour, local stitched slot)."""
        n_cams = len(self.layout.points)
        jobs = []
        for f in range(self.n_frames):
            for cam in range(n_cams):
                z, c = self.ptrs(cam, f)
                remote = 1 if self.layout.rank_of(cam) != self.rank else 0     # PCS_B200_JOB_REMOTE_FRAME
                jobs.append((cam, z, c, stitched_frames[f].slot_ptr(cam), None, None, remote))
        return jobs

    def barrier(self):
        """All ranks have finished reading each other's frames (call before overwriting them)."""
        self.handle.barrier()


class ShardedMerge:
    """Voxel merge of a multi-camera cloud that is NEVER assembled: sharded by voxel-key range before the exchange
    (SURVEY s8(e)).  Every rank holds only its own cameras' records; per frame

        hist (my points per z plane) -> barrier -> plan (all ranks derive the same equal-population z cuts from
        everybody's histogram, read over NVLink) -> scatter (all-to-all: each record goes to the inbox of the rank
        that owns its slab, peer stores from the binning kernel) -> barrier -> merge of my inbox

    and slab r of the merged grid stays on rank r (``out``, ``count``).  The slabs concatenated in rank order are the
    single-GPU merge of all cameras, bit for bit.  The exchange buffers (histogram, inbox cursor, inbox) live in
    one symmetric-memory allocation per frame slot; torch symmetric memory is used for the allocation, the
    rendezvous and the device-side barrier only.  Nothing in a frame returns to the host.
    """

    def __init__(self, ctx, rank, world, device, n_total, leaf_mm, n_slots=1, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        import pointcloud_stitching_b200 as pcs
        self.ctx, self.rank, self.world, self.leaf = ctx, rank, world, leaf_mm
        self.capacity = int(n_total)               # worst case: every point of the cloud falls into one slab
        self.zbins = pcs.lib.pcs_b200_shard_zbins(leaf_mm)
        if self.zbins < 0:
            raise ValueError("bad leaf")
        al = lambda b: (b + 255) & ~255            # noqa: E731
        self.o_cursor, self.o_zhist = 0, 256
        self.o_inbox = self.o_zhist + al(self.zbins * 4)
        self.slot_bytes = self.o_inbox + al(self.capacity * RECORD_BYTES)
        self.n_slots = n_slots
        self.raw = symm.empty(self.slot_bytes * n_slots, dtype=torch.uint8, device=device)
        self.raw.zero_()
        self.handle = symm.rendezvous(self.raw, group if group is not None else dist.group.WORLD)
        bases = list(self.handle.buffer_ptrs)
        assert bases[rank] == self.raw.data_ptr()
        self.peers = []
        for k in range(n_slots):
            p = pcs.ShardPeers()
            p.n_ranks, p.rank, p.capacity_records = world, rank, self.capacity
            for r in range(world):
                b = bases[r] + k * self.slot_bytes
                p.cursor_dev[r], p.zhist_dev[r], p.inbox_dev[r] = b + self.o_cursor, b + self.o_zhist, b + self.o_inbox
            self.peers.append(p)
        # local, per slot: the cuts, the plane -> rank table, the error flag, the voxel count
        self.splits = torch.zeros(n_slots, world + 1, dtype=torch.int32, device=device)
        self.zslab = torch.zeros(n_slots, (self.zbins + 15) & ~15, dtype=torch.uint8, device=device)
        self.err = torch.zeros(n_slots, dtype=torch.int32, device=device)
        self.count = torch.zeros(n_slots, dtype=torch.int32, device=device)

    def local_ptrs(self, slot):
        base = self.raw.data_ptr() + slot * self.slot_bytes
        return base + self.o_cursor, base + self.o_zhist, base + self.o_inbox

    def barrier(self):
        self.handle.barrier()

    def run(self, slot, records_ptr, n_own, out, cuda_stream=0):
        """Enqueue one frame on the CURRENT torch stream (``cuda_stream`` is its raw handle): ``out`` (int16, capacity
        ``capacity * 5``) receives this rank's slab, ``self.count[slot]`` its voxel count, ``self.splits[slot]`` the cuts."""
        cursor, zhist, inbox = self.local_ptrs(slot)
        p = self.peers[slot]
        self.ctx.shard_hist_dev(records_ptr, n_own, self.leaf, zhist, cursor, cuda_stream)
        self.handle.barrier()
        self.ctx.shard_plan_dev(p, self.leaf, self.splits[slot].data_ptr(), self.zslab[slot].data_ptr(), cuda_stream)
        self.ctx.shard_scatter_dev(records_ptr, n_own, self.leaf, self.zslab[slot].data_ptr(), p, self.err[slot].data_ptr(),
                                   cuda_stream)
        self.handle.barrier()
        self.ctx.voxel_merge_counted_async_dev(inbox, self.capacity, cursor, self.leaf, out.data_ptr(),
                                               self.count[slot].data_ptr(), cuda_stream)


def sharded_voxel_merge(ctx, records_ptr, n, leaf_mm, rank, world, out, cuda_stream=0, gather=True, group=None):
    """Voxel merge of a stitched cloud that every rank holds, sharded by z-slab.

    Every rank computes the same equal-population cuts of the grid along z from its own copy of the
    records (no communication), merges slab ``rank`` only -- 1/world of the sort work -- into
    ``out`` (int16 tensor, capacity n * 5) and, with ``gather``, the slabs are exchanged so that
    every rank ends with the whole merged cloud in ascending (kz, ky, kx) order, exactly what the
    single-GPU ``voxel_merge_dev`` returns.  Returns (voxels in ``out``, voxels of this rank's slab).
    Without ``gather`` each rank keeps its slab at the start of ``out`` (a z-range of the grid per GPU).
    """
    if world == 1:
        mine = ctx.voxel_merge_dev(records_ptr, n, leaf_mm, out.data_ptr(), cuda_stream)
        return mine, mine
    splits, pts = ctx.voxel_slab_plan_dev(records_ptr, n, leaf_mm, world, cuda_stream)
    if not gather:
        mine = ctx.voxel_merge_slab_dev(records_ptr, n, leaf_mm, splits[rank], splits[rank + 1], out.data_ptr(),
                                        cuda_stream)
        return mine, mine
    # A slab cannot produce more voxels than it has points, and the plan gives every rank all the slab
    # populations: each rank merges into its block of a [world x cap] staging buffer, ONE all-gather
    # moves the blocks (the first 8 bytes of a block carry its voxel count), and the slabs are then
    # packed back to back in slab order.
    blk = (4 + max(pts) * 5 + 7) // 8 * 8        # int16 per block: [int64 voxel count][records], 16-byte multiple
    stage = torch.empty(world * blk, dtype=torch.int16, device=out.device)
    block = stage[rank * blk:(rank + 1) * blk]
    mine = ctx.voxel_merge_slab_dev(records_ptr, n, leaf_mm, splits[rank], splits[rank + 1], block[4:].data_ptr(),
                                    cuda_stream)
    block[:4].view(torch.int64)[0] = mine
    dist.all_gather_into_tensor(stage.view(torch.uint8), block.view(torch.uint8), group=group)   # NCCL has no int16
    counts = stage.view(world, blk)[:, :4].contiguous().view(torch.int64).reshape(-1).tolist()
    start = 0
    for r in range(world):
        c = int(counts[r])
        if c:
            out[start * 5:(start + c) * 5].copy_(stage[r * blk + 4:r * blk + 4 + c * 5])
        start += c
    return start, mine
