#!/usr/bin/env python
"""Per-kernel figures of one or more `ncu --set full` captures as JSON (launches of a kernel are averaged):

    python tools/ncu_kernels_json.py out.json capture1.ncu-rep [capture2.ncu-rep ...]
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

KEEP = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__occupancy_limit_registers": "occ_limit_regs", "launch__occupancy_limit_shared_mem": "occ_limit_smem",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "smsp__inst_executed.sum": "warp_instructions",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle"}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    agg = collections.OrderedDict()
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            a = agg.setdefault(name, {"launches": 0, "source": os.path.basename(rep)})
            a["launches"] += 1
            for m, key in KEEP.items():
                if m not in hdr:
                    continue
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if units[i] in SCALE and ("bytes" in m or "time" in m):
                    v *= SCALE[units[i]]
                a[key] = a.get(key, 0.0) + v
    for a in agg.values():
        for k in list(a):
            if k not in ("launches", "source"):
                a[k] /= a["launches"]
        if "duration_us" in a and "dram_read_bytes" in a:
            a["dram_GBps"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["duration_us"] / 1e3
    json.dump({"kernels": agg, "how": "ncu --set full --clock-control none --import-source on, launches of a kernel averaged; "
               "cold caches, serialised: shares agree with the CUDA-event timings, absolutes are a little higher"},
              open(out, "w"), indent=1)
    for k, a in agg.items():
        print("%-28s %8.1f us  %6.0f GB/s  regs %3d  warps %4.1f%%" % (k[:28], a.get("duration_us", 0), a.get("dram_GBps", 0),
                                                                    a.get("registers", 0), a.get("warps_active_pct", 0)))


if __name__ == "__main__":
    main()
