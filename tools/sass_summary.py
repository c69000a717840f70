#!/usr/bin/env python
"""Per-kernel SASS opcode census of libpcs_b200.so (cuobjdump -sass): the mnemonics that prove what a kernel is
made of -- UBLKCP (TMA bulk copies), SYNCS (mbarrier), FFMA2 / FMUL2 (packed fp32x2), ATOMS / ATOMG / RED
(atomics), LDG / STG / LDS / STS widths, REDUX, MATCH, VOTE -- as a markdown table.

    python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FFMA", "IMAD", "LDG", "LDG.128", "STG", "STG.128", "LDS", "STS", "STS.128",
        "ATOMS", "ATOMG", "RED", "REDUX", "MATCH", "VOTE", "SHFL", "BAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pointcloud_stitching_b200", "libpcs_b200.so")
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur, arch = None, set()
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if base in ("LDG", "STG", "STS", "LDS") and ".128" in op:
                cur[base + ".128"] += 1
    print("# SASS opcode census of `%s` (%s; %d kernels)\n" % (os.path.relpath(lib, ROOT), ", ".join(sorted(arch)), len(kernels)))
    print("| kernel | instr | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    tot = collections.Counter()
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", name).replace("pcs::", "")
        print("| `%s` | %d | " % (short[:60], c["total"]) + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")
        tot.update(c)
    print("| **all** | %d | " % tot["total"] + " | ".join(str(tot[k]) if tot[k] else "" for k in COLS) + " |")


if __name__ == "__main__":
    main()
