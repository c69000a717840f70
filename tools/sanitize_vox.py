#!/usr/bin/env python
"""Small voxel-merge / slab / PLY run for compute-sanitizer (memcheck, racecheck):

    compute-sanitizer --tool racecheck python tools/sanitize_vox.py

164 000 points = 41 sort tiles (three look-back groups), one skewed voxel, checked against the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import pointcloud_stitching_b200 as pcs  # noqa: E402

rng = np.random.default_rng(3)
n = 164000
rec = rng.integers(-32768, 32767, (n, 5)).astype(np.int16)
rec[:, :3] = rng.normal(0, 900, (n, 3)).clip(-32768, 32767).astype(np.int16)
rec[rng.random(n) < 0.2, :3] = (7, -3, 1500)
rec[:, 4] &= 0xFF
R = oracle.restatement()
want = R.voxel_merge(rec, 10)
ctx = pcs.Context(device=0, max_streams=1)
d = torch.from_numpy(rec.reshape(-1)).cuda()
out = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
cs = torch.cuda.current_stream().cuda_stream
nv = ctx.voxel_merge_dev(d.data_ptr(), n, 10, out.data_ptr(), cs)
torch.cuda.synchronize()
assert nv == len(want) and np.array_equal(out[: nv * 5].cpu().numpy().reshape(-1, 5), want)
splits, pts = ctx.voxel_slab_plan_dev(d.data_ptr(), n, 10, 3, cs)
got = []
for r in range(3):
    k = ctx.voxel_merge_slab_dev(d.data_ptr(), n, 10, splits[r], splits[r + 1], out.data_ptr(), cs)
    torch.cuda.synchronize()
    got.append(out[: k * 5].cpu().numpy().reshape(-1, 5).copy())
assert np.array_equal(np.concatenate(got), want)
# slab partition + bitmap ranking (voxel_variant 4), the enqueue-only merge, and the all-to-all kernels of the sharded
# multi-GPU merge with three "ranks" on this GPU
c4 = pcs.Context(device=0, max_streams=1, voxel_variant=4)
nv4 = c4.voxel_merge_dev(d.data_ptr(), n, 10, out.data_ptr(), cs)
torch.cuda.synchronize()
assert nv4 == len(want) and np.array_equal(out[: nv4 * 5].cpu().numpy().reshape(-1, 5), want)
cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
ctx.voxel_merge_async_dev(d.data_ptr(), n, 10, out.data_ptr(), cnt.data_ptr(), cs)
torch.cuda.synchronize()
assert int(cnt[0].item()) == len(want)
zbins = pcs.lib.pcs_b200_shard_zbins(10)
third = (n // 3) & ~7
parts = [(0, third), (third, 2 * third), (2 * third, n)]
zh = [torch.zeros(zbins, dtype=torch.int32, device="cuda") for _ in parts]
cur = [torch.zeros(1, dtype=torch.int32, device="cuda") for _ in parts]
inbox = [torch.zeros(n * 5 + 8, dtype=torch.int16, device="cuda") for _ in parts]
peers = pcs.ShardPeers()
peers.n_ranks, peers.rank, peers.capacity_records = 3, 0, n
for r in range(3):
    peers.inbox_dev[r], peers.cursor_dev[r], peers.zhist_dev[r] = inbox[r].data_ptr(), cur[r].data_ptr(), zh[r].data_ptr()
for r, (a, b) in enumerate(parts):
    ctx.shard_hist_dev(d.data_ptr() + a * 10, b - a, 10, zh[r].data_ptr(), cur[r].data_ptr(), cs)
sp = torch.zeros(4, dtype=torch.int32, device="cuda")
zslab = torch.zeros((zbins + 15) & ~15, dtype=torch.uint8, device="cuda")
err = torch.zeros(1, dtype=torch.int32, device="cuda")
ctx.shard_plan_dev(peers, 10, sp.data_ptr(), zslab.data_ptr(), cs)
for r, (a, b) in enumerate(parts):
    ctx.shard_scatter_dev(d.data_ptr() + a * 10, b - a, 10, zslab.data_ptr(), peers, err.data_ptr(), cs)
got = []
for r in range(3):
    ctx.voxel_merge_counted_async_dev(inbox[r].data_ptr(), n, cur[r].data_ptr(), 10, out.data_ptr(), cnt.data_ptr() + 4 * r, cs)
    torch.cuda.synchronize()
    got.append(out[: int(cnt[r].item()) * 5].cpu().numpy().reshape(-1, 5).copy())
assert int(err.item()) == 0 and np.array_equal(np.concatenate(got), want)
cloud = torch.randn(1000 * 8, device="cuda")
rows = torch.zeros(1000 * 15, dtype=torch.uint8, device="cuda")
ctx.cloud_to_ply_rows_dev(cloud.data_ptr(), 1000, rows.data_ptr(), cs)
torch.cuda.synchronize()
print("sanitize_vox: OK", nv, pts)
