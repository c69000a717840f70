#!/usr/bin/env python
"""bench.py -- Mpoints/s stitched on synthetic 1280x720 depth+RGB streams.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path (depth deprojection + 4x4 transform + colour
attach + int16 pack, fused kernel K1) over one batch of STREAMS x FRAMES synthetic
frames per GPU (default 8 x 8 = 64 frames = 59.0 Mpoints, 885 MB of algorithmic
traffic -- seven times the 126 MB L2, so every step streams from HBM).  With N > 1
(torchrun, one rank per GPU) every rank runs its own 8 streams (weak scaling) and
the packed records are all-gathered over NVLink into the stitched buffer.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events, max over
ranks) with inputs resident in HBM; `e2e` goes through the reference-facing C ABI
call (pcs_b200_send_xyzrgb: host buffers in, host camera buffer out), host<->device
copies inside the timed region.  `--impl reference` times the reference's own CPU
implementation (oracle/_ref: src/pcs-camera-optimized.cpp compiled unmodified) on
the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1280, 720
NPTS = W * H
ALG_BYTES_PER_POINT = 15          # 2 (z16) + 3 (RGB8) + 10 (record), SURVEY s8(d)
METRIC = "Mpoints/sec stitched (1280x720xN cams)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="camera streams per GPU")
    ap.add_argument("--frames", type=int, default=8, help="frames per stream per step")
    ap.add_argument("--variant", type=int, default=0, help="kernel_variant: 0 auto, 1 direct, 2 pipelined")
    ap.add_argument("--tex", default="baseline", choices=["baseline", "aligned"],
                    help="depth->colour extrinsics: 15 mm baseline (D435-like) or identity")
    ap.add_argument("--exchange", default="pull", choices=["pull", "fused", "nccl"],
                    help="N > 1: pull = every rank runs K1 over all cameras, reading the peers' raw frames over NVLink "
                         "(5 B/pt on the link); fused = K1 stores every tile of records to all peers (10 B/pt); "
                         "nccl = K1 then all-gather")
    ap.add_argument("--e2e-depth", type=int, default=2, choices=[1, 2], help="frames in flight per camera in the e2e leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r01_k1_ncu_summary.json")
    try:
        return json.load(open(p)).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_frames(streams, frames, rank):
    from pointcloud_stitching_b200 import synth
    d = np.empty((streams, frames, H, W), np.uint16)
    c = np.empty((streams, frames, H, W * 3), np.uint8)
    for s in range(streams):
        for f in range(frames):
            d[s, f] = synth.depth_frame(W, H, rank * streams + s, f)
            c[s, f] = synth.color_frame(W, H, rank * streams + s, f)
    return d, c


# ------------------------------------------------------------------------------------
def cpu_reference(threads, sample_frames, tex):
    """The reference CPU path on `sample_frames` 1280x720 frames with `threads` OpenMP threads:
    stub rs2::pointcloud::calculate (oracle restatement of librealsense, threaded the same way)
    + the reference's own sendXYZRGBPointcloud (-m).  Returns dict."""
    import oracle
    from pointcloud_stitching_b200 import synth
    R = oracle.restatement()
    RC = oracle.ref_camera()
    trans = synth.D2C_BASELINE if tex == "baseline" else (0.0, 0.0, 0.0)
    cal = oracle.make_calib(W, H, translation=trans)
    z, col = synth.depth_frame(W, H, 0, 0), synth.color_frame(W, H, 0, 0)
    t_calc = []
    xyz = uv = None
    for _ in range(3 + sample_frames):
        t0 = time.perf_counter()
        xyz, uv = R.deproject(cal, z, threads)
        t_calc.append((time.perf_counter() - t0) * 1e3)
    t_calc = t_calc[3:]
    if RC is not None:
        kind = "reference"
        RC.time_send(xyz, uv, col, W, H, 3, W * 3, synth.TF_CAMERA, 3, threads=threads)
        t_pack = list(RC.time_send(xyz, uv, col, W, H, 3, W * 3, synth.TF_CAMERA, sample_frames, threads=threads))
    else:
        kind = "port"
        t_pack = []
        for _ in range(3 + sample_frames):
            t0 = time.perf_counter()
            R.send(xyz, uv, col, W, H, 3, W * 3, synth.TF_CAMERA)
            t_pack.append((time.perf_counter() - t0) * 1e3)
        t_pack = t_pack[3:]
        threads = 1
    calc, pack = float(np.median(t_calc)), float(np.median(t_pack))
    return {"value": NPTS / ((calc + pack) * 1e-3) / 1e6, "unit": "Mpoints/s", "cores": threads, "kind": kind,
            "sample": "%d frames 1280x720, median per frame: deproject (oracle restatement of rs2::pointcloud::"
                      "calculate) %.3f ms + sendXYZRGBPointcloud -m -t %d %.3f ms" % (sample_frames, calc, threads, pack),
            "pack_only_mpoints_s": NPTS / (pack * 1e-3) / 1e6, "deproject_ms": calc, "pack_ms": pack,
            "host_cores": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = args.streams                      # one frame per stream: a bounded sample of the step
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_reference(threads, per_step, args.tex)
        if i >= args.warmup:
            vals.append(info["value"])
    v = float(np.median(vals))
    ms = per_step * NPTS / (v * 1e6) * 1e3
    info["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpoints/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d streams x 1280x720 depth+RGB8, d2c=%s; each step is a bounded sample of "
                                   "%d frames (one per stream) on the host cores" % (args.streams, args.tex, per_step)},
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import pointcloud_stitching_b200 as pcs
    from pointcloud_stitching_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, F = args.streams, args.frames
    trans = synth.D2C_BASELINE if args.tex == "baseline" else (0.0, 0.0, 0.0)

    # ABI streams S..2S-1 mirror 0..S-1: the e2e leg double-buffers every camera (two frames in flight)
    ctx = pcs.Context(device=local, max_streams=2 * S, kernel_variant=args.variant)
    for s in range(S):
        cam = rank * S + s
        for k in (s, S + s):
            ctx.set_stream(k, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=trans))

    d_np, c_np = make_frames(S, F, rank)
    d_dev = torch.from_numpy(d_np.view(np.int16)).cuda()
    c_dev = torch.from_numpy(c_np).cuda()
    # stitched buffers, one per frame index: [pad 12][int32 bytes][world x S cameras x records]
    from pointcloud_stitching_b200 import multigpu
    slot = S * NPTS * 10
    layout = multigpu.StitchLayout([NPTS] * (world * S), world)
    exchange = "none" if world == 1 else args.exchange
    sset = fset = None
    if exchange == "pull":
        try:
            fset = multigpu.SymmetricFrameSet(layout, rank, torch.device("cuda", local), W, H, F)
            for s in range(S):
                for f in range(F):
                    fset.upload(rank * S + s, f, d_np[s, f], c_np[s, f])
            # every rank computes every camera: one context whose stream ids are the camera indices
            pctx = pcs.Context(device=local, max_streams=world * S, kernel_variant=args.variant)
            for cam in range(world * S):
                pctx.set_stream(cam, pcs.stream_desc(W, H, tf=synth.TF_STITCH[cam % 8], translation=trans))
        except Exception as e:
            if rank == 0:
                print("bench: symmetric memory unavailable (%r); falling back to --exchange nccl" % (e,), file=sys.stderr)
            exchange, fset = "nccl", None
    if exchange == "fused":
        try:
            sset = multigpu.SymmetricStitchedSet(layout, rank, torch.device("cuda", local), F)
            stitched = sset.frames
        except Exception as e:  # no symmetric memory on this box: say so and use the NCCL baseline
            if rank == 0:
                print("bench: symmetric memory unavailable (%r); falling back to --exchange nccl" % (e,), file=sys.stderr)
            exchange = "nccl"
    if sset is None:
        stitched = [multigpu.StitchedBuffer(layout, rank, torch.device("cuda", local)) for _ in range(F)]
    rec_views = [st.payload for st in stitched]

    def job(s, f):
        return (s, d_dev[s, f].data_ptr(), c_dev[s, f].data_ptr(), stitched[f].slot_ptr(rank * S + s))

    cs = torch.cuda.current_stream()
    all_jobs = [job(s, f) for f in range(F) for s in range(S)]
    if world == 1:
        batches = [ctx.batch(all_jobs)]
    elif exchange == "pull":
        batches = [pctx.batch(fset.pull_jobs(stitched))]
    elif exchange == "fused":
        batches = [ctx.batch_fanout(all_jobs, sset.local_base, sset.nbytes, sset.peer_bases)]
    else:
        batches = [ctx.batch([job(s, f) for s in range(S)]) for f in range(F)]
        comm = torch.cuda.Stream()
    launches_per_step = sum(b.launches for b in batches)

    def step():
        if world == 1:
            batches[0].run(cs.cuda_stream)
            return
        if exchange == "pull":
            batches[0].run(cs.cuda_stream)
            fset.barrier()       # nobody is still reading this rank's frames
            return
        if exchange == "fused":
            batches[0].run(cs.cuda_stream)
            sset.barrier()
            return
        for f in range(F):
            batches[f].run(cs.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(cs)
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(rec_views[f], rec_views[f][rank * slot:(rank + 1) * slot])
        cs.wait_stream(comm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # A step is ~0.2 ms, so the timed region is a few milliseconds, shorter than nvidia-smi's
    # sampling period: keep the GPU busy with the same work for ~0.15 s first (a couple of clock
    # samples under this very load), then time K steps back to back; the sampler runs across both.
    # (A much longer ramp -- 0.5 s -- runs this 1 kW part into sw_power_cap at ~1870 MHz, -8 %.)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ramp_steps = max(1, int(0.15 / 0.0002)) if world == 1 else 30
    for _ in range(ramp_steps):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs)
    for _ in range(args.steps):
        step()
    e1.record(cs)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    pts_step = world * S * F * NPTS
    value = pts_step / (ms_step * 1e-3) / 1e6

    # kernel-only timing of K1 without any exchange (for N > 1 the step above also moves the
    # records over NVLink); this is what the HBM roofline refers to
    local_batches = batches if world == 1 else [ctx.batch(all_jobs)]
    for b in local_batches:
        b.run(cs.cuda_stream)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(cs)
    for _ in range(args.steps):
        for b in local_batches:
            b.run(cs.cuda_stream)
    k1.record(cs)
    torch.cuda.synchronize()
    ms_kernel_step = k0.elapsed_time(k1) / args.steps
    launches_per_step_local = sum(b.launches for b in local_batches)
    # roofline: the dominant kernel's launch time over the timed region itself at N = 1 (a step IS
    # one k1_pipe launch there); for N > 1 the step also holds the exchange, so the K1-only loop
    launch_ms = (ms_step if world == 1 else ms_kernel_step) / launches_per_step_local
    peak, peak_src = measured_peak()
    alg_bytes_launch = ALG_BYTES_PER_POINT * S * F * NPTS / launches_per_step_local
    achieved = alg_bytes_launch / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "peak_source": peak_src, "kernel": "k1_pipe" if args.variant != 1 else "k1_direct",
                "algorithmic_bytes_per_launch": alg_bytes_launch, "launch_ms": launch_ms}

    # ---- end to end through the reference-facing C-ABI call, host buffers -------------
    e2e = None
    if not args.no_e2e:
        hz = [[ctx.host_alloc(NPTS * 2, np.uint16) for _ in range(F)] for _ in range(S)]
        hc = [[ctx.host_alloc(NPTS * 3, np.uint8) for _ in range(F)] for _ in range(S)]
        hb = [[ctx.new_camera_buffer(pinned=True) for _ in range(S)] for _ in range(2)]
        for s in range(S):
            for f in range(F):
                hz[s][f][:] = d_np[s, f].reshape(-1)
                hc[s][f][:] = c_np[s, f].reshape(-1)

        def e2e_step():
            total = 0
            if args.e2e_depth == 1:
                for f in range(F):
                    for s in range(S):
                        ctx.send_begin(s, hz[s][f], hc[s][f], hb[0][s], True)
                    for s in range(S):
                        total += ctx.send_end(s)
                return total
            # software pipeline: frame f of every camera is in flight while frame f-1 drains
            for f in range(F):
                slot = f & 1
                if f >= 2:
                    for s in range(S):
                        total += ctx.send_end(slot * S + s)
                for s in range(S):
                    ctx.send_begin(slot * S + s, hz[s][f], hc[s][f], hb[slot][s], True)
            for f in range(max(0, F - 2), F):
                for s in range(S):
                    total += ctx.send_end((f & 1) * S + s)
            return total

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 10))
        for _ in range(n_e2e):
            got = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert got == S * F * NPTS * 10
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * S * F * NPTS * n_e2e / dt / 1e6, "unit": "Mpoints/s",
               "h2d_bytes_per_step": S * F * NPTS * 5, "d2h_bytes_per_step": S * F * NPTS * 10,
               "api": "pcs_b200_send_xyzrgb_begin/_end (host z16+RGB8 in, reference camera buffer out), "
                      "%d cameras x %d frame(s) in flight, pinned host buffers" % (S, args.e2e_depth), "steps": n_e2e}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(os.cpu_count() or 1, 40, args.tex)
        cpu1 = cpu_reference(1, 20, args.tex)
        cpu["single_thread_mpoints_s"] = cpu1["value"]
        cpu["single_thread_pack_only_mpoints_s"] = cpu1["pack_only_mpoints_s"]

    link_div = 2 if exchange == "pull" else 1
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d streams/GPU x %d frames x 1280x720 z16 depth + RGB8, fused deproject+transform+"
                                   "colour+pack (K1)%s; depth->colour extrinsics: %s" % (
                                       S, F, "" if world == 1 else {
                                           "pull": " on every rank over ALL cameras' frames: the peers' raw z16+RGB8 tiles are the "
                                                   "kernel's own TMA loads from NVLink peer memory (all-gather fused into the "
                                                   "compute kernel's input pipeline, 5 B/pt on the link)",
                                           "fused": " + all-gather of the packed records fused into the kernel (peer TMA stores "
                                                    "over NVLink, 10 B/pt on the link)",
                                           "nccl": " + in-place NCCL all-gather of the packed records"}[exchange],
                                       "15 mm baseline" if args.tex == "baseline" else "identity"),
                       "streams_per_gpu": S, "frames_per_step": F, "points_per_step": pts_step,
                       "l2": "working set %.0f MB per step per GPU >> 126 MB L2 (no flush needed)" % (
                           ALG_BYTES_PER_POINT * S * F * NPTS / 1e6),
                       "kernel_variant": args.variant, "tex": args.tex, "exchange": exchange},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
            "kernel_only": {"ms_per_step": ms_kernel_step, "mpoints_s_per_gpu": S * F * NPTS / (ms_kernel_step * 1e-3) / 1e6},
            "nvlink": None if world == 1 else {
                "recv_bytes_per_gpu_per_step": (world - 1) * slot * F // link_div,
                "recv_GBps_per_gpu": (world - 1) * slot * F / link_div / (ms_step * 1e-3) / 1e9,
                "note": "every GPU must take in (N-1)/N of the stitched cloud -- as 10 B/pt records (fused, nccl) or as "
                        "5 B/pt raw frames it deprojects itself (pull): this link rate, not HBM, bounds N > 1 "
                        "(B200_PROFILING.md: 770 GB/s measured peer copy per direction)"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
