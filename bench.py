#!/usr/bin/env python
"""bench.py -- Mpoints/s stitched on synthetic 1280x720 depth+RGB streams.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--tex baseline|aligned|rotated|color1080p|rotated1080p]

Headline ("value"): one "step" = one pass of the hot path (depth deprojection + 4x4 transform + colour
attach + int16 pack, fused kernel K1) over one batch of STREAMS x FRAMES synthetic frames per GPU
(default 8 x 8 = 64 frames = 59.0 Mpoints, 885 MB of algorithmic traffic -- seven times the 126 MB L2,
so every step streams from HBM).  With N > 1 (torchrun, one rank per GPU) every rank owns 8 streams
(weak scaling) and ends each step holding the stitched buffer of ALL N x 8 cameras (exchange fused
into K1's input pipeline over NVLink); the stitched bytes are verified once per run on every rank
("stitched_check").

The same JSON line carries, under "configs", BASELINE.json's other single-box configurations as
self-checking measurements: c3 (4 x 1280x720 on one GPU: K1 -> in-place concat -> 10 mm voxel merge) and
c5 (20 x 848x480 on the run's N GPUs: K1 -> exchange -> sharded voxel merge), each with its own
roofline and cpu_baseline.

Prints ONE JSON line (rank 0).  `value` is device-timed (CUDA events, max over ranks) with inputs
resident in HBM; `e2e` goes through the reference-facing C ABI with HOST buffers (host<->device copies
inside the timed region).  `--impl reference` times the reference's own CPU implementation
(oracle/_ref: src/pcs-camera-optimized.cpp compiled unmodified) on the host cores, on the same
config; that arm never loads the product library.
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
LEAF_MM = 10
TEX_CHOICES = ["baseline", "aligned", "rotated", "color1080p", "rotated1080p"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8, help="camera streams per GPU")
    ap.add_argument("--frames", type=int, default=8, help="frames per stream per step")
    ap.add_argument("--variant", type=int, default=0, help="kernel_variant: 0 auto, 1 direct, 2 pipelined")
    ap.add_argument("--tex", default="baseline", choices=TEX_CHOICES,
                    help="depth->colour calibration: 15 mm baseline (D435-like), identity, 15 mm + small rotation, "
                         "or 15 mm + 1920x1080 colour (what the reference records, src/pcs-camera-grab-frames.cpp:69-70)")
    ap.add_argument("--exchange", default="pull", choices=["pull", "fused", "nccl"],
                    help="N > 1: pull = every rank runs K1 over all cameras, reading the peers' raw frames over NVLink "
                         "(5 B/pt on the link); fused = K1 stores every tile of records to all peers (10 B/pt); "
                         "nccl = K1 then all-gather")
    ap.add_argument("--configs", default="c3,c5", help="extra BASELINE configs measured into the same line ('none' to skip)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained K1 leg (0 to skip)")
    ap.add_argument("--e2e-slots", type=int, default=2, choices=[1, 2, 3, 4], help="stitched frames in flight in the e2e leg")
    ap.add_argument("--merge-lanes", type=int, default=8, help="c3 / c5 at N = 1: stitched frames merged concurrently")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the stitched-bytes verification")
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


def ncu_traffic(name):
    """dram read+write bytes per launch of a kernel from a committed ncu capture (not measured in this
    run: ncu cannot run inside the timed process).  Returns (bytes or None, source label)."""
    for rnd in ("r02", "r01"):
        rel = os.path.join("profiles", "%s_%s_ncu_summary.json" % (rnd, name))
        try:
            v = json.load(open(os.path.join(ROOT, rel))).get("dram_bytes_per_launch")
            if v is not None:
                return v, rel + " (ncu --set full capture of the same launch, committed)"
        except Exception:
            pass
    return None, None


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
        return self

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


# ---- the synthetic calibration of a --tex choice (shared by both arms and the oracle) -----------
def tex_calibration(tex):
    """dict(cw, ch, translation, rotation) -- colour geometry and depth->colour extrinsics."""
    from pointcloud_stitching_b200 import synth
    cal = {"cw": W, "ch": H, "translation": synth.D2C_BASELINE, "rotation": None}
    if tex == "aligned":
        cal["translation"] = (0.0, 0.0, 0.0)
    elif tex == "rotated":
        cal["rotation"] = synth.D2C_ROTATION_SMALL
    elif tex == "color1080p":
        cal["cw"], cal["ch"] = 1920, 1080
    elif tex == "rotated1080p":      # the reference's recording geometry (src/pcs-camera-grab-frames.cpp:69-70) on a real D4xx
        cal["cw"], cal["ch"], cal["rotation"] = 1920, 1080, synth.D2C_ROTATION_SMALL
    return cal


def tex_label(tex):
    return {"baseline": "15 mm baseline", "aligned": "identity", "rotated": "15 mm baseline + 0.3 deg rotation",
            "color1080p": "15 mm baseline, 1920x1080 colour",
            "rotated1080p": "15 mm baseline + 0.3 deg rotation, 1920x1080 colour"}[tex]


def headline_config(args, world):
    """`config` of the JSON line: identical for the b200 arm and the reference arm of the same command."""
    S, F = args.streams, args.frames
    cal = tex_calibration(args.tex)
    return {"workload": "%d streams/GPU x %d frames x 1280x720 z16 depth + %dx%d RGB8 per step: fused deproject + "
                        "transform + colour + int16 pack, stitched in camera order; depth->colour: %s" % (
                            S, F, cal["cw"], cal["ch"], tex_label(args.tex)),
            "streams_per_gpu": S, "frames_per_step": F, "points_per_step": world * S * F * NPTS, "tex": args.tex,
            "l2": "working set %.0f MB per step per GPU >> 126 MB L2 (distinct inputs, no flush needed)" % (
                ALG_BYTES_PER_POINT * S * F * NPTS / 1e6)}


def make_frames(streams, frames, rank, cw=W, ch=H, w=W, h=H, first_cam=None):
    from pointcloud_stitching_b200 import synth
    first = rank * streams if first_cam is None else first_cam
    d = np.empty((streams, frames, h, w), np.uint16)
    c = np.empty((streams, frames, ch, cw * 3), np.uint8)
    for s in range(streams):
        for f in range(frames):
            d[s, f] = synth.depth_frame(w, h, first + s, f)
            c[s, f] = synth.color_frame(cw, ch, first + s, f)
    return d, c


# ------------------------------------------------------------------------------------
def cpu_reference(threads, frames, tex, d_np=None, c_np=None):
    """The reference CPU path over `frames` DISTINCT 1280x720 frames with `threads` OpenMP threads:
    stub rs2::pointcloud::calculate (oracle restatement of librealsense, threaded the same way) + the
    reference's own sendXYZRGBPointcloud (-m -t threads, oracle/_ref).  Returns dict."""
    import oracle
    from pointcloud_stitching_b200 import synth
    R = oracle.restatement()
    RC = oracle.ref_camera()
    cal_d = tex_calibration(tex)
    cw, ch = cal_d["cw"], cal_d["ch"]
    cal = oracle.make_calib(W, H, cw, ch, translation=cal_d["translation"], rotation=cal_d["rotation"])
    if d_np is None:
        n_src = min(frames, 8)
        d_np, c_np = make_frames(1, n_src, 0, cw, ch)
    d_flat, c_flat = d_np.reshape(-1, H, W), c_np.reshape(-1, ch, cw * 3)
    kind = "reference" if RC is not None else "port"
    if RC is None:
        threads = 1
    t_calc, t_pack = [], []
    for i in range(2 + frames):
        z, col = d_flat[i % len(d_flat)], c_flat[i % len(c_flat)]
        t0 = time.perf_counter()
        xyz, uv = R.deproject(cal, z, threads)
        t1 = time.perf_counter()
        if RC is not None:
            ms = float(RC.time_send(xyz, uv, col, cw, ch, 3, cw * 3, synth.TF_CAMERA, 1, threads=threads)[0])
        else:
            t2 = time.perf_counter()
            R.send(xyz, uv, col, cw, ch, 3, cw * 3, synth.TF_CAMERA)
            ms = (time.perf_counter() - t2) * 1e3
        if i >= 2:
            t_calc.append((t1 - t0) * 1e3)
            t_pack.append(ms)
    calc, pack = float(np.sum(t_calc)), float(np.sum(t_pack))
    return {"value": frames * NPTS / ((calc + pack) * 1e-3) / 1e6, "unit": "Mpoints/s", "cores": threads, "kind": kind,
            "sample": "%d distinct frames 1280x720 (%s), per frame: deproject (oracle restatement of rs2::pointcloud::"
                      "calculate) %.3f ms + sendXYZRGBPointcloud -m -t %d %.3f ms" % (
                          frames, tex, calc / frames, threads, pack / frames),
            "pack_only_mpoints_s": frames * NPTS / (pack * 1e-3) / 1e6, "deproject_ms": calc / frames,
            "pack_ms": pack / frames, "total_ms": calc + pack, "host_cores": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on this box's host cores, same config.  Loads
    oracle/ only -- never the product library (pointcloud_stitching_b200.lib is lazy and untouched here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    threads = os.cpu_count() or 1
    S, F = args.streams, args.frames
    cal = tex_calibration(args.tex)
    d_np, c_np = make_frames(S, F, 0, cal["cw"], cal["ch"])
    frames_step_full = world * S * F
    frames_sample = S * F            # one GPU's share of the step, all S x F distinct frames, every step
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_reference(threads, frames_sample, args.tex, d_np, c_np)
        if i >= args.warmup:
            vals.append(info["value"])
    v = float(np.median(vals))
    ms = frames_step_full * NPTS / (v * 1e6) * 1e3
    info["value"] = v
    info["sample"] = ("every step: %d of the step's %d frames (all %d x %d distinct frames of one GPU's share), all %d host "
                      "threads; ms_per_step is scaled to the full step; last step: %s" % (
                          frames_sample, frames_step_full, S, F, threads, info["sample"]))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpoints/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(args, world),
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "product_library_loaded": "libpcs_b200" in open("/proc/self/maps").read()}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------
class Timer:
    """CUDA-event timing on torch's current stream, max over ranks."""

    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, iters, warm=3):
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / iters


def stream_descs(pcs, synth, tex, n_cams, w=W, h=H):
    cal = tex_calibration(tex)
    cw, ch = (cal["cw"], cal["ch"]) if (w, h) == (W, H) else (w, h)
    return [pcs.stream_desc(w, h, cw, ch, tf=synth.TF_STITCH[cam % 8], translation=cal["translation"],
                            rotation=cal["rotation"]) for cam in range(n_cams)]


def oracle_records(tex, cam, z, col, w=W, h=H):
    """CPU oracle of K1 for one frame (the checker; never timed as product)."""
    import oracle
    from pointcloud_stitching_b200 import synth
    cal = tex_calibration(tex)
    cw, ch = (cal["cw"], cal["ch"]) if (w, h) == (W, H) else (w, h)
    R = oracle.restatement()
    c = oracle.make_calib(w, h, cw, ch, translation=cal["translation"], rotation=cal["rotation"])
    return R.frame(c, z, col, 3, cw * 3, synth.TF_STITCH[cam % 8])


# ---- config #3: 4 x 1280x720 on one GPU, K1 -> in-place concat -> voxel merge ------------------
def bench_voxel_config(name, torch, dist, pcs, synth, multigpu, args, rank, world, local, timer, peak):
    """c3: 4 x 1280x720, one GPU.  c5: 20 x 848x480 on `world` GPUs (BASELINE.json configs #3 / #5).
    Step = `frames` stitched frames: per frame one K1 launch writes every camera's slot of the stitched buffer
    (the concat costs nothing) and the stitched records are merged on a 10 mm voxel grid."""
    if name == "c3":
        cams, w, h, frames = 4, 1280, 720, 8
        if world > 1:
            return None      # a single-GPU configuration
    else:
        cams, w, h, frames = 20, 848, 480, 8
    npts = w * h
    dev = torch.device("cuda", local)
    cs = torch.cuda.current_stream().cuda_stream
    n = cams * npts
    layout = multigpu.StitchLayout([npts] * cams, world)
    ctx = pcs.Context(device=local, max_streams=cams, kernel_variant=args.variant)
    stride = ((w * 3 + 15) // 16) * 16
    for cam in range(cams):
        d = pcs.stream_desc(w, h, tf=synth.TF_STITCH[cam % 8], translation=synth.D2C_BASELINE)
        d.color_stride = stride
        ctx.set_stream(cam, d)
    keep = []
    sm = None
    if world > 1:
        # every rank deprojects ITS cameras only; the records are never assembled: they go straight to the rank that
        # owns their z slab of the voxel grid (multigpu.ShardedMerge)
        mine = layout.cams_of[rank]
        own = [torch.zeros(max(1, len(mine)) * npts * 5, dtype=torch.int16, device=dev) for _ in range(frames)]
        batches = []
        for f in range(frames):
            jobs = []
            for i, cam in enumerate(mine):
                col = np.zeros((h, stride), np.uint8)
                col[:, :w * 3] = synth.color_frame(w, h, cam, f)
                z = torch.from_numpy(synth.depth_frame(w, h, cam, f).view(np.int16)).to(dev)
                c = torch.from_numpy(col).to(dev)
                keep.append((z, c))
                jobs.append((cam, z.data_ptr(), c.data_ptr(), own[f].data_ptr() + i * npts * 10))
            batches.append(ctx.batch(jobs) if jobs else None)
        # `xl` exchange lanes: consecutive frames run hist / all-to-all / merge concurrently on their own streams (own
        # merge context, own symmetric buffers and barrier pads each) -- at N = 8 a slab is ~1 M points and a lone
        # frame is all launch and barrier latency
        xl = max(1, min(args.merge_lanes, frames))
        xctx = [ctx] + [pcs.Context(device=local, max_streams=1) for _ in range(xl - 1)]
        xstreams = [torch.cuda.current_stream()] + [torch.cuda.Stream() for _ in range(xl - 1)]
        sms = [multigpu.ShardedMerge(xctx[k], rank, world, dev, n, LEAF_MM, n_slots=(frames + xl - 1) // xl) for k in range(xl)]
        sm = sms[0]
        stitched = None
    else:
        stitched = [multigpu.StitchedBuffer(layout, rank, dev) for _ in range(frames)]
        batches = []
        for f in range(frames):
            jobs = []
            for cam in range(cams):
                col = np.zeros((h, stride), np.uint8)
                col[:, :w * 3] = synth.color_frame(w, h, cam, f)
                z = torch.from_numpy(synth.depth_frame(w, h, cam, f).view(np.int16)).to(dev)
                c = torch.from_numpy(col).to(dev)
                keep.append((z, c))
                jobs.append((cam, z.data_ptr(), c.data_ptr(), stitched[f].slot_ptr(cam)))
            batches.append(ctx.batch(jobs))
    outs = [torch.zeros(n * 5, dtype=torch.int16, device=dev) for _ in range(frames)]
    nv = [0] * frames
    counts = torch.zeros(frames, dtype=torch.int32, device=dev)

    def k1_only():
        for f in range(frames):
            if batches[f] is not None:
                batches[f].run(cs)

    # world == 1: enqueue-only merges (box, key layout, pass count and voxel count stay on the device; the counts are
    # read once, after the step's last frame) on `lanes` CUDA streams, one merge context (scratch + plan slot) each:
    # consecutive stitched frames are merged concurrently, which fills the latency-bound phases of the sort
    lanes = args.merge_lanes if world == 1 else 1
    mctx = [ctx] + [pcs.Context(device=local, max_streams=1) for _ in range(lanes - 1)]
    mstreams = [torch.cuda.current_stream()] + [torch.cuda.Stream() for _ in range(lanes - 1)]

    def merge_only():
        if world == 1:
            main = mstreams[0]
            for s2 in mstreams[1:]:
                s2.wait_stream(main)              # the frames' records (K1) are complete on the main stream
            for f in range(frames):
                k = f % lanes
                mctx[k].voxel_merge_async_dev(stitched[f].payload.data_ptr(), n, LEAF_MM, outs[f].data_ptr(),
                                              counts.data_ptr() + 4 * f, mstreams[k].cuda_stream)
            for s2 in mstreams[1:]:
                main.wait_stream(s2)
            return
        # per frame: hist -> barrier -> plan -> all-to-all -> barrier -> merge of my slab; nothing returns to the host
        main = xstreams[0]
        for s2 in xstreams[1:]:
            s2.wait_stream(main)
        for f in range(frames):
            k = f % xl
            with torch.cuda.stream(xstreams[k]):
                sms[k].run(f // xl, own[f].data_ptr(), len(layout.cams_of[rank]) * npts, outs[f], xstreams[k].cuda_stream)
        for s2 in xstreams[1:]:
            main.wait_stream(s2)

    def step():
        k1_only()
        merge_only()

    iters = max(3, min(args.steps, 10))
    ms_step = timer.run(step, iters)
    ms_k1 = timer.run(k1_only, iters)
    ms_merge = timer.run(merge_only, iters)
    if world == 1:
        nv = [int(v) for v in counts.cpu().tolist()]
    else:
        nv = [int(sms[f % xl].count[f // xl].item()) for f in range(frames)]
        if any(int(x.err.abs().sum().item()) != 0 for x in sms):
            raise RuntimeError("an inbox overflowed")
    if min(nv) < 0:
        raise RuntimeError("voxel merge reported status %d" % min(nv))
    # voxels over all ranks (each rank keeps its z-slab of the grid)
    nv_local = torch.tensor([float(sum(nv))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(nv_local)
    nv_total = int(nv_local.item())
    pts_step = n * frames
    # roofline of the merge, SURVEY s8(d): lower bound 10 B/pt read + 10 B/voxel written
    alg = 10.0 * pts_step / world + 10.0 * sum(nv)
    achieved = alg / (ms_merge * 1e-3) / 1e9
    res = {"workload": "%d cams x %dx%d, %d stitched frames per step, K1 -> %s -> voxel-grid merge %d mm" % (
               cams, w, h, frames, "in-place concat" if world == 1 else
               "records sharded by z slab of the voxel grid BEFORE the exchange (per-rank z histogram, identical cuts on every "
               "GPU, all-to-all as peer stores over NVLink: 10 B/pt x (N-1)/N, each record crosses once); slab r stays on GPU r",
               LEAF_MM),
           "n_gpus": world, "points_per_step": pts_step, "voxels_per_step": nv_total,
           "value": pts_step / (ms_step * 1e-3) / 1e6, "unit": "Mpoints/s", "ms_per_step": ms_step,
           "k1_ms_per_step": ms_k1, "merge_ms_per_step": ms_merge, "merge_ms_per_frame": ms_merge / frames,
           "merge_lanes": lanes if world == 1 else xl,
           "gpu_launches_per_step": None,
           "roofline": {"bound": "hbm", "kernel": "voxel merge (sw_keys_hist + sw_pass x P + sw_reduce; sw_pass dominant)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "algorithmic_bytes_per_step_per_gpu": alg, "traffic": None,
                        "note": "bytes per SURVEY s8(d): 10 B/point read + 10 B/voxel written, over the merge's device time"}}
    # ---- self-check, frame 0: this rank's voxels against the CPU oracle of the same inputs
    check = "skipped"
    if not args.no_check:
        import oracle
        R = oracle.restatement()
        recs = []
        for cam in range(cams):
            recs.append(oracle_records("baseline", cam, synth.depth_frame(w, h, cam, 0), synth.color_frame(w, h, cam, 0), w, h))
        t0 = time.perf_counter()
        want_vox = R.voxel_merge(np.concatenate(recs), LEAF_MM)
        t_vox = time.perf_counter() - t0
        if world == 1:
            got_st = stitched[0].payload.cpu().numpy().view(np.int16).reshape(-1, 5)
            ok = np.array_equal(got_st, np.concatenate(recs))
            got = outs[0][: nv[0] * 5].cpu().numpy().reshape(-1, 5)
            ok = ok and nv[0] == len(want_vox) and np.array_equal(got, want_vox)
        else:
            # this rank's slab of the oracle's merge of ALL cameras, cut where the GPUs cut
            splits = sm.splits[0].cpu().numpy()
            kz = np.floor_divide(want_vox[:, 2].astype(np.int32), LEAF_MM)
            sel = want_vox[(kz >= splits[rank]) & (kz < splits[rank + 1])]
            got = outs[0][: nv[0] * 5].cpu().numpy().reshape(-1, 5)
            ok = nv[0] == len(sel) and np.array_equal(got, sel)
            allcuts = [torch.zeros_like(sm.splits[0]) for _ in range(world)]
            dist.all_gather(allcuts, sm.splits[0].contiguous())
            ok = ok and all(bool(torch.equal(c, allcuts[0])) for c in allcuts)
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        check = "ok" if flag.item() == 1.0 else "MISMATCH"
        if rank == 0 and not args.no_cpu and world == 1:
            # CPU baseline on the same frame: the reference's concat loop (sendStitchToUnity, compiled from
            # src/pcs-multicamera-client.cpp when oracle/_ref is present) + the oracle's voxel merge (the
            # reference has no voxel grid of its own), one thread as in the reference's stitcher
            t0 = time.perf_counter()
            R.concat([r.reshape(-1) for r in recs], 1)
            t_cat = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": n / (t_cat + t_vox) / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
                                   "sample": "1 stitched frame (%d points): concat loop of sendStitchToUnity "
                                             "(src/pcs-multicamera-client.cpp:385-392, restated) %.1f ms + oracle voxel merge "
                                             "(qsort, own spec) %.1f ms" % (n, t_cat * 1e3, t_vox * 1e3)}
    res["check"] = check
    for b in batches:
        if b is not None:
            b.close()
    for c in mctx[1:]:
        c.close()
    if world > 1:
        for c in xctx[1:]:
            c.close()
    ctx.close()
    return res


# ------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import pointcloud_stitching_b200 as pcs
    from pointcloud_stitching_b200 import multigpu, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    timer = Timer(torch, dist, world)
    S, F = args.streams, args.frames
    cal = tex_calibration(args.tex)
    cw, ch = cal["cw"], cal["ch"]
    descs = stream_descs(pcs, synth, args.tex, world * S)

    # ABI streams S..2S-1 mirror 0..S-1: the e2e leg double-buffers every camera (two frames in flight)
    ctx = pcs.Context(device=local, max_streams=2 * S, kernel_variant=args.variant)
    for s in range(S):
        for k in (s, S + s):
            ctx.set_stream(k, descs[rank * S + s])

    d_np, c_np = make_frames(S, F, rank, cw, ch)
    d_dev = torch.from_numpy(d_np.view(np.int16)).cuda()
    c_dev = torch.from_numpy(c_np).cuda()
    # stitched buffers, one per frame index: [pad 12][int32 bytes][world x S cameras x records]
    slot = S * NPTS * 10
    layout = multigpu.StitchLayout([NPTS] * (world * S), world)
    exchange = "none" if world == 1 else args.exchange
    sset = fset = pctx = None
    if exchange == "pull":
        try:
            fset = multigpu.SymmetricFrameSet(layout, rank, torch.device("cuda", local), W, H, F, stride=cw * 3,
                                              color_height=ch)
            for s in range(S):
                for f in range(F):
                    fset.upload(rank * S + s, f, d_np[s, f], c_np[s, f])
            # every rank computes every camera: one context whose stream ids are the camera indices
            pctx = pcs.Context(device=local, max_streams=world * S, kernel_variant=args.variant)
            for cam in range(world * S):
                pctx.set_stream(cam, descs[cam])
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

    barrier = timer.barrier
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # A step is ~0.2 ms, so the timed region is a few milliseconds, shorter than nvidia-smi's
    # sampling period: keep the GPU busy with the same work for ~0.15 s first (a couple of clock
    # samples under this very load), then time K steps back to back; the sampler runs across both.
    # (A much longer ramp runs this 1 kW part into sw_power_cap at ~1870 MHz: that is the "sustained"
    # figure reported beside this burst figure.)
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

    # ---- the stitched bytes of this run, verified on every rank -------------------------------------
    stitched_check = "skipped"
    if not args.no_check:
        ok = True
        # (a) frame 0 of one camera per rank (own cameras AND pulled / received peers' cameras) against a K1-local
        #     recompute from regenerated inputs; (b) on rank 0 one peer camera against the CPU oracle
        cams_to_check = sorted({r * S for r in range(world)} | {rank * S + S - 1})
        cctx = pcs.Context(device=local, max_streams=1, kernel_variant=args.variant)
        for cam in cams_to_check:
            z, col = synth.depth_frame(W, H, cam, 0), synth.color_frame(cw, ch, cam, 0)
            cctx.set_stream(0, descs[cam])
            dz, dc = torch.from_numpy(z.view(np.int16)).cuda(), torch.from_numpy(col).cuda()
            pay = torch.zeros(NPTS * 10, dtype=torch.uint8, device="cuda")
            b = cctx.batch([(0, dz.data_ptr(), dc.data_ptr(), pay.data_ptr())])
            b.run(cs.cuda_stream)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(pay, stitched[0].slot(cam)))
            b.close()
            if rank == 0 and cam == cams_to_check[-1 if world > 1 else 0]:
                want = oracle_records(args.tex, cam, z, col)
                ok = ok and np.array_equal(pay.cpu().numpy().view(np.int16).reshape(-1, 5), want)
        hdr = int(np.frombuffer(stitched[0].wire_bytes()[:4].cpu().numpy().tobytes(), np.int32)[0])
        ok = ok and hdr == world * S * NPTS * 10
        cctx.close()
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        stitched_check = "ok" if flag.item() == 1.0 else "MISMATCH"

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
    # colour bytes per depth pixel: the rows the taps can land in.  Same-size frames: every row, 3 B/px.  A colour frame
    # of another size behind a pure x baseline: depth row y taps colour row floor((y - ppy) / fy * cfy + cppy + .5) whatever
    # the depth (the library proves it per stream and reads only those rows), so the rows in between are never needed
    if args.tex in ("color1080p", "rotated1080p"):      # (rotated: the same count of rows, sheared along x)
        tapped = len({int(np.floor((y - (H - 1) / 2) / (W / 2) * (cw / 2) + (ch - 1) / 2 + 0.5)) for y in range(H)})
    else:
        tapped = ch
    color_bytes = tapped * cw * 3 / NPTS
    alg_bpp = 2 + color_bytes + 10
    alg_bytes_launch = alg_bpp * S * F * NPTS / launches_per_step_local
    achieved = alg_bytes_launch / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic("k1") if args.tex == "baseline" else (None, None)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "k1_pipe" if args.variant != 1 else "k1_direct",
                "algorithmic_bytes_per_point": alg_bpp, "algorithmic_bytes_per_launch": alg_bytes_launch, "launch_ms": launch_ms,
                "colour_rows_tapped": tapped, "colour_rows": ch,
                "frac_if_whole_colour_frame_counted": (2 + ch * cw * 3 / NPTS + 10) * S * F * NPTS / launches_per_step_local
                / (launch_ms * 1e-3) / 1e9 / peak}

    # ---- sustained: the same K1 step back to back for >= 2 s (power-capped clocks), beside the burst figure
    sustained = None
    if args.sustained_seconds > 0:
        sam2 = ClockSampler(local)
        if rank == 0:
            sam2.start()
        n_sus = max(args.steps, int(args.sustained_seconds / (ms_kernel_step * 1e-3)) + 1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(cs)
        for _ in range(n_sus):
            for b in local_batches:
                b.run(cs.cuda_stream)
        s1.record(cs)
        torch.cuda.synchronize()
        sus_ms = s0.elapsed_time(s1)
        sus_clocks = sam2.stop() if rank == 0 else None
        sus_ach = alg_bpp * S * F * NPTS * n_sus / (sus_ms * 1e-3) / 1e9
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_sus, "value": S * F * NPTS * n_sus / (sus_ms * 1e-3) / 1e6,
                     "unit": "Mpoints/s per GPU (K1 only, no exchange)", "achieved": sus_ach, "frac": sus_ach / peak,
                     "clocks": sus_clocks}

    # ---- end to end through the reference-facing C-ABI call, host buffers -------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, pcs, synth, ctx, descs, d_np, c_np, rank, world, local, timer)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference(os.cpu_count() or 1, 40, args.tex, d_np, c_np)
        cpu1 = cpu_reference(1, 12, args.tex, d_np, c_np)
        cpu["single_thread_mpoints_s"] = cpu1["value"]
        cpu["single_thread_pack_only_mpoints_s"] = cpu1["pack_only_mpoints_s"]

    # ---- BASELINE configs #3 and #5 into the same line ------------------------------------------------
    configs = {}
    want = [] if args.configs in ("none", "") else [c.strip() for c in args.configs.split(",")]
    for name in want:
        try:
            r = bench_voxel_config(name, torch, dist, pcs, synth, multigpu, args, rank, world, local, timer, peak)
        except Exception as e:  # a failing extra config must not lose the headline; it is reported as failed
            r = {"error": repr(e), "check": "ERROR"}
        if r is not None:
            configs[name] = r

    link_div = 2 if exchange == "pull" else 1
    bad = stitched_check == "MISMATCH" or any(c.get("check") in ("MISMATCH", "ERROR") for c in configs.values())
    e2e_bad = e2e is not None and e2e.get("check") == "MISMATCH"
    if world > 1:
        fl = torch.tensor([1.0 if e2e_bad else 0.0], device="cuda")
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
        e2e_bad = fl.item() > 0
    bad = bad or e2e_bad
    if rank == 0:
        nvlink = None
        if world > 1:
            gbps = (world - 1) * slot * F / link_div / (ms_step * 1e-3) / 1e9
            nvlink = {"recv_bytes_per_gpu_per_step": (world - 1) * slot * F // link_div, "recv_GBps_per_gpu": gbps,
                      "peer_copy_peak_GBps": 770.0, "frac_of_peer_copy_peak": gbps / 770.0,
                      "note": "every GPU must take in (N-1)/N of the stitched cloud -- as 10 B/pt records (fused, nccl) or as "
                              "5 B/pt raw frames it deprojects itself (pull): this link rate, not HBM, bounds N > 1 "
                              "(B200_PROFILING.md: 770 GB/s measured peer copy per direction)"}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(args, world),
            "impl_detail": {"kernel_variant": args.variant, "exchange": exchange, "exchange_note": {
                "none": "single GPU: K1 writes every camera's slot of the stitched buffer in place",
                "pull": "K1 on every rank over ALL cameras' frames: the peers' raw z16+RGB8 tiles are the kernel's own TMA "
                        "loads from NVLink peer memory (all-gather fused into the compute kernel's input pipeline, 5 B/pt)",
                "fused": "all-gather of the packed records fused into K1 (peer TMA stores over NVLink, 10 B/pt)",
                "nccl": "K1 then in-place NCCL all-gather of the packed records"}[exchange]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "sustained": sustained, "cpu_baseline": cpu, "stitched_check": stitched_check,
            "kernel_only": {"ms_per_step": ms_kernel_step, "mpoints_s_per_gpu": S * F * NPTS / (ms_kernel_step * 1e-3) / 1e6},
            "nvlink": nvlink, "configs": configs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if bad:
        sys.exit("bench.py: output verification FAILED (stitched_check=%s, configs=%s)" % (
            stitched_check, {k: v.get("check") for k, v in configs.items()}))


def run_e2e(args, torch, dist, pcs, synth, ctx, descs, d_np, c_np, rank, world, local, timer):
    """The same metric through the reference-facing C ABI with HOST buffers on both sides: pinned z16 + RGB8
    frames of this GPU's S cameras in, the reference's STITCHED buffer `[int32][cam0 records][cam1 records]...`
    out (pcs_b200_stitch_frames_begin/_end: per camera and CUDA stream H2D, k1_pipe into the camera's slot of the
    device-resident stitched buffer, D2H into its place of the stitched host buffer), frames in flight on several slots;
    host<->device copies inside the timed region.
    Beside it: the per-camera call (pcs_b200_send_xyzrgb_begin/_end, S separate camera buffers) and a plain
    cudaMemcpy probe moving the same bytes (what the PCIe link / host memory allow with no kernel at all)."""
    S, F = args.streams, args.frames
    ch, cwb = c_np.shape[2], c_np.shape[3]
    hz = [[ctx.host_alloc(NPTS * 2, np.uint16) for _ in range(F)] for _ in range(S)]
    hc = [[ctx.host_alloc(ch * cwb, np.uint8) for _ in range(F)] for _ in range(S)]
    hb = [[ctx.new_camera_buffer(pinned=True) for _ in range(S)] for _ in range(2)]
    NS = args.e2e_slots
    hs = [ctx.host_alloc(4 + S * NPTS * 10, np.uint8) for _ in range(NS)]
    for s in range(S):
        for f in range(F):
            hz[s][f][:] = d_np[s, f].reshape(-1)
            hc[s][f][:] = c_np[s, f].reshape(-1)
    cams = list(range(S))

    def stitched_step():
        total = 0
        for f in range(F):
            slot = f % NS
            if f >= NS:
                total += ctx.stitch_frames_end(slot)
            ctx.stitch_frames_begin(slot, cams, [hz[s][f] for s in range(S)], [hc[s][f] for s in range(S)], hs[slot])
        for f in range(max(0, F - NS), F):
            total += ctx.stitch_frames_end(f % NS)
        return total

    def camera_step():
        total = 0
        for f in range(F):      # software pipeline: frame f of every camera is in flight while frame f-1 drains
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

    # plain copies of the same bytes in the same structure (per frame slot and camera one stream: frame up, records down
    # into their place of the stitched host buffer), no kernel
    tz = [[torch.from_numpy(hz[s][f].view(np.int16)) for f in range(F)] for s in range(S)]
    tc = [[torch.from_numpy(hc[s][f]) for f in range(F)] for s in range(S)]
    ts = [torch.from_numpy(hs[k]) for k in range(NS)]
    dz = [[torch.empty(NPTS, dtype=torch.int16, device="cuda") for _ in range(S)] for _ in range(NS)]
    dc = [[torch.empty(ch * cwb, dtype=torch.uint8, device="cuda") for _ in range(S)] for _ in range(NS)]
    ds = [torch.empty(4 + S * NPTS * 10, dtype=torch.uint8, device="cuda") for _ in range(NS)]
    streams = [[torch.cuda.Stream() for _ in range(S)] for _ in range(NS)]

    def probe_step():
        for f in range(F):
            k = f % NS
            if f >= NS:
                for st in streams[k]:
                    st.synchronize()
            for s in range(S):
                with torch.cuda.stream(streams[k][s]):
                    dz[k][s].copy_(tz[s][f], non_blocking=True)
                    dc[k][s].copy_(tc[s][f], non_blocking=True)
                    lo, hi = 4 + s * NPTS * 10, 4 + (s + 1) * NPTS * 10
                    ts[k][lo:hi].copy_(ds[k][lo:hi], non_blocking=True)
        for k in range(NS):
            for st in streams[k]:
                st.synchronize()
        return S * F * NPTS * 10

    n_e2e = max(3, min(args.steps, 10))

    def wall(fn):
        fn()
        timer.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            got = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert got == S * F * NPTS * 10
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * S * F * NPTS * n_e2e / dt / 1e6

    v_stitched = wall(stitched_step)
    # the bytes the call returned for the last frame: header + this GPU's cameras against a device-side K1 of the same frame
    last = hs[(F - 1) % NS]
    hdr = int(last[:4].view(np.int32)[0])
    chk = torch.zeros(S * NPTS * 10, dtype=torch.uint8, device="cuda")
    zt = [torch.from_numpy(d_np[s, F - 1].view(np.int16)).cuda() for s in range(S)]
    ct = [torch.from_numpy(c_np[s, F - 1]).cuda() for s in range(S)]
    b = ctx.batch([(s, zt[s].data_ptr(), ct[s].data_ptr(), chk.data_ptr() + s * NPTS * 10) for s in range(S)])
    b.run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ok = hdr == S * NPTS * 10 and bool(np.array_equal(last[4:], chk.cpu().numpy()))
    b.close()
    v_camera = wall(camera_step)
    v_probe = wall(probe_step)
    h2d, d2h = S * F * (NPTS * 2 + ch * cwb), S * F * NPTS * 10 + 4 * F
    return {"value": v_stitched, "unit": "Mpoints/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "api": "pcs_b200_stitch_frames_begin/_end (host z16+RGB8 of %d cameras in, the reference's stitched buffer "
                   "[int32][records] out; every camera on its own stream: frame up, k1_pipe into its slot of the device-resident "
                   "stitched buffer, records down into their place), %d frames in flight, pinned host buffers" % (S, NS),
            "steps": n_e2e, "check": "ok" if ok else "MISMATCH",
            "per_camera_api": {"value": v_camera, "api": "pcs_b200_send_xyzrgb_begin/_end (%d separate camera buffers, "
                               "2 frames in flight)" % S},
            "pcie_probe": {"value": v_probe, "what": "plain cudaMemcpyAsync of the same bytes per step in the same structure (one stream "
                           "per frame slot and camera: frame up, records down), no kernel", "d2h_GBps": v_probe * 1e6 * 10 / 1e9 / world,
                           "h2d_GBps": v_probe * 1e6 * (2 + ch * cwb / NPTS) / 1e9 / world},
            "frac_of_pcie": v_stitched / v_probe}


if __name__ == "__main__":
    main()
