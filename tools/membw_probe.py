#!/usr/bin/env python
"""HBM micro-probe: read-only, write-only, copy and a 1:2 read:write mix (the K1 traffic shape),
CUDA-event timed with torch ops.  Context for roofline.frac: what can this part sustain for a
write-heavy stream?  Writes gpurun_out/membw_probe.json."""
import json
import os
import torch

def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(iters):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

n = 1 << 30  # bytes
a = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 255)
b = torch.empty(n, dtype=torch.uint8, device="cuda")
a32, b32 = a.view(torch.int32), b.view(torch.int32)
res = {}
ms = timed(lambda: b32.fill_(7)); res["write_only_GBps"] = n / ms / 1e6
ms = timed(lambda: b32.copy_(a32)); res["copy_GBps(read+write)"] = 2 * n / ms / 1e6
ms = timed(lambda: a32.sum()); res["read_only_GBps"] = n / ms / 1e6
# 1 read : 2 write -- expand N/3 int32 into 2N/3 (two copies)
k = (n // 4) // 3
src = a32[:k]
dst = b32[: 2 * k].view(2, k)
ms = timed(lambda: dst.copy_(src.unsqueeze(0).expand(2, k))); res["mix_1r2w_GBps"] = 3 * k * 4 / ms / 1e6
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/membw_probe.json", "w"), indent=1)
print(json.dumps(res))
