#!/usr/bin/env python
"""Turn an ncu capture (gpurun_out/*.ncu-rep) and an ncu launch list (launches.csv) into the
small tracked artefacts under profiles/:

    python tools/ncu_summary.py gpurun_out/prof_k1_pipe.ncu-rep gpurun_out/launches.csv r01

writes profiles/<tag>_k1_ncu_summary.json (read by bench.py for roofline.traffic),
profiles/<tag>_k1_stalls.md and profiles/<tag>_launches.md.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
name = sys.argv[4] if len(sys.argv) > 4 else "k1"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True, check=True).stdout


def num(s):
    return float(s.replace(",", ""))


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
kernels = []
for r in raw[2:]:
    k = {"kernel": r[hdr.index("Kernel Name")]}
    for m in KEEP:
        if m in hdr:
            i = hdr.index(m)
            v = num(r[i]) if r[i] not in ("", "n/a") else None
            if v is not None and units[i] in UNIT_SCALE and ("bytes" in m or "time" in m):
                v *= UNIT_SCALE[units[i]]
            k[m] = v
    kernels.append(k)
k0 = kernels[0]
dram = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in kernels) / len(kernels)
summary = {"source": os.path.basename(rep), "command": "ncu --set full --clock-control none --import-source on "
           "-k regex:k1_pipe -s 3 -c 2 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e",
           "kernel": k0["kernel"], "launches_captured": len(kernels),
           "dram_bytes_per_launch": dram,
           "dram_read_bytes_per_launch": sum(k["dram__bytes_read.sum"] for k in kernels) / len(kernels),
           "dram_write_bytes_per_launch": sum(k["dram__bytes_write.sum"] for k in kernels) / len(kernels),
           "duration_us": sum(k["gpu__time_duration.sum"] for k in kernels) / len(kernels),
           "metrics_first_launch": k0}
json.dump(summary, open(os.path.join(out, "%s_%s_ncu_summary.json" % (tag, name)), "w"), indent=1)

# ---- per-instruction stalls
src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
hi = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
h = src[hi]
rows = []
for r in src[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if r and r[0].startswith("0x") and len(r) > h.index("stall_wait"):
        rows.append(r)
tot = sum(int(r[h.index("# Samples")]) for r in rows)
cols = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_math", "stall_not_selected", "stall_selected"]
with open(os.path.join(out, "%s_%s_stalls.md" % (tag, name)), "w") as f:
    f.write("# %s: warp-stall samples by SASS instruction, %s\n\n" % (tag, k0["kernel"]))
    f.write("%d instructions, %d samples (first captured launch).  Totals by reason:\n\n" % (len(rows), tot))
    for c in cols:
        f.write("* %s: %d\n" % (c, sum(int(r[h.index(c)]) for r in rows)))
    f.write("\n| samples | executed | long_sb | short_sb | wait | mio | math | SASS |\n|---|---|---|---|---|---|---|---|\n")
    for r in sorted(rows, key=lambda r: -int(r[h.index("# Samples")]))[:40]:
        f.write("| %s | %s | %s | %s | %s | %s | %s | `%s` |\n" % (
            r[h.index("# Samples")], r[h.index("Instructions Executed")], r[h.index("stall_long_sb")],
            r[h.index("stall_short_sb")], r[h.index("stall_wait")], r[h.index("stall_mio")], r[h.index("stall_math")],
            r[h.index("Source")].strip()))
    ops = collections.Counter()
    for r in rows:
        op = r[h.index("Source")].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        ops[op.split(".")[0]] += int(r[h.index("Instructions Executed")])
    f.write("\nExecuted warp-instructions by opcode: " + ", ".join("%s %d" % kv for kv in ops.most_common(24)) + "\n")

# ---- launch list
agg = collections.OrderedDict()
for r in csv.reader(open(launches)):
    if not r or not r[0].isdigit():
        continue
    v = num(r[-1]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[-2], 1.0)
    a = agg.setdefault(r[4].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
with open(os.path.join(out, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: every kernel launch of `python bench.py --steps 2 --warmup 3 --no-cpu` under\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)\n\n" % tag)
    f.write("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f%% | %.2f |\n" % (n, c, t, 100 * t / total, t / c))
    f.write("\nThe timed region of `value` holds only `k1_pipe` launches (one per step); `k1_direct` and the\n"
            "fill kernels belong to the e2e leg (one frame per call through pcs_b200_send_xyzrgb) and to setup.\n")
print(json.dumps({k: summary[k] for k in ("dram_bytes_per_launch", "duration_us")}))
