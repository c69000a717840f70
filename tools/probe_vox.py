"""Timing probe for the voxel merge: whole-merge time on 14.7 M uniform random points under the tuning knobs
PCS_SW_PROBE (needs a library built with `make -C pointcloud_stitching_b200/csrc -B PTXAS=-DPCS_SW_PROBES`;
bit 0: skip the look-back, 1: skip the stores, 2: skip the ranking -- results are then WRONG, only the time
means anything), PCS_SW_BALLOT (0 = MATCH.ANY ranking), PCS_SW_POLL_NS and VV (voxel_variant).
Results: profiles/r01_voxel.md."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import pointcloud_stitching_b200 as pcs
rng = np.random.default_rng(1)
n = 14745600
rec = torch.from_numpy(rng.integers(-6000, 6000, (n, 5)).astype(np.int16)).cuda()
out = torch.zeros(n * 5, dtype=torch.int16, device="cuda")
ctx = pcs.Context(device=0, max_streams=1, voxel_variant=int(os.environ.get("VV", "0")))
cs = torch.cuda.current_stream().cuda_stream
def run():
    try:
        return ctx.voxel_merge_dev(rec.data_ptr(), n, 10, out.data_ptr(), cs)
    except pcs.PcsError as e:
        return -1
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): nv = run()
e1.record(); torch.cuda.synchronize()
print("vv", os.environ.get("VV", "0"), os.environ.get("PCS_SW_PROBE", "0"), os.environ.get("PCS_SW_BALLOT", "1"), "ms", e0.elapsed_time(e1) / 10, "nv", nv)
