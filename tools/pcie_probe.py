#!/usr/bin/env python
"""How fast can host<->device copies of the e2e leg's byte counts go, by copy structure?  (measurement aid)

Per stitched frame: 8 cameras x (1.84 MB z16 + 2.76 MB RGB8) up, 73.7 MB of records down.  Variants differ in
how the copies are cut and which streams carry them.  Prints one JSON line (GB/s per direction, Mpoints/s)."""
import json
import time

import torch

S, NPTS, F = 8, 1280 * 720, 16


def main():
    hz = [[torch.empty(NPTS, dtype=torch.int16).pin_memory() for _ in range(4)] for _ in range(S)]
    hc = [[torch.empty(NPTS * 3, dtype=torch.uint8).pin_memory() for _ in range(4)] for _ in range(S)]
    hs = [torch.empty(S * NPTS * 10 + 16, dtype=torch.uint8).pin_memory() for _ in range(4)]
    hsc = [[torch.empty(NPTS * 10, dtype=torch.uint8).pin_memory() for _ in range(S)] for _ in range(4)]
    dz = [[torch.empty(NPTS, dtype=torch.int16, device="cuda") for _ in range(S)] for _ in range(4)]
    dc = [[torch.empty(NPTS * 3, dtype=torch.uint8, device="cuda") for _ in range(S)] for _ in range(4)]
    ds = [torch.empty(S * NPTS * 10 + 16, dtype=torch.uint8, device="cuda") for _ in range(4)]
    res = {}

    def run(name, slots, d2h_mode, up_streams=1):
        st = [[torch.cuda.Stream() for _ in range(max(S, 2))] for _ in range(slots)]

        def frame(k):
            main_s = st[k][0]
            if up_streams == 1:
                with torch.cuda.stream(main_s):
                    for s in range(S):
                        dz[k][s].copy_(hz[s][k], non_blocking=True)
                        dc[k][s].copy_(hc[s][k], non_blocking=True)
            else:
                for s in range(S):
                    with torch.cuda.stream(st[k][s]):
                        dz[k][s].copy_(hz[s][k], non_blocking=True)
                        dc[k][s].copy_(hc[s][k], non_blocking=True)
                for s in range(1, S):
                    main_s.wait_stream(st[k][s])
            if d2h_mode == "one":
                with torch.cuda.stream(main_s):
                    hs[k].copy_(ds[k], non_blocking=True)
            elif d2h_mode == "chunks_same":
                with torch.cuda.stream(main_s):
                    for s in range(S):
                        hsc[k][s].copy_(ds[k][s * NPTS * 10:(s + 1) * NPTS * 10], non_blocking=True)
            elif d2h_mode == "chunks_streams":
                for s in range(S):
                    st[k][s].wait_stream(main_s)
                    with torch.cuda.stream(st[k][s]):
                        hsc[k][s].copy_(ds[k][s * NPTS * 10:(s + 1) * NPTS * 10], non_blocking=True)
            elif d2h_mode == "two_halves":
                half = S * NPTS * 5
                with torch.cuda.stream(main_s):
                    hs[k][:half].copy_(ds[k][:half], non_blocking=True)
                st[k][1].wait_stream(main_s)
                with torch.cuda.stream(st[k][1]):
                    hs[k][half:2 * half].copy_(ds[k][half:2 * half], non_blocking=True)

        def sync(k):
            for x in st[k]:
                x.synchronize()

        def step():
            for f in range(F):
                k = f % slots
                if f >= slots:
                    sync(k)
                frame(k)
            for k in range(slots):
                sync(k)

        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        pts = S * NPTS * F
        res[name] = {"mpoints_s": pts / dt / 1e6, "d2h_GBps": pts * 10 / dt / 1e9, "h2d_GBps": pts * 5 / dt / 1e9}

    run("one_d2h_2slots", 2, "one")
    run("one_d2h_3slots", 3, "one")
    run("chunks_same_stream_2slots", 2, "chunks_same")
    run("chunks_8streams_2slots", 2, "chunks_streams")
    run("two_halves_2slots", 2, "two_halves")
    run("one_d2h_2slots_up8streams", 2, "one", up_streams=8)
    run("chunks_8streams_2slots_up8streams", 2, "chunks_streams", up_streams=8)
    run("chunks_8streams_4slots_up8streams", 4, "chunks_streams", up_streams=8)
    # one direction at a time
    s0 = torch.cuda.Stream()
    for name, fn in (("d2h_only", lambda: hs[0].copy_(ds[0], non_blocking=True)),
                     ("h2d_only", lambda: [dc[0][s].copy_(hc[s][0], non_blocking=True) for s in range(S)])):
        with torch.cuda.stream(s0):
            fn()
            s0.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            s0.synchronize()
            dt = (time.perf_counter() - t0) / 10
        nbytes = hs[0].numel() if name == "d2h_only" else S * NPTS * 3
        res[name] = {"GBps": nbytes / dt / 1e9}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
