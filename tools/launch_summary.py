#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total, share, average per kernel."""
import collections
import csv
import io
import sys


def main(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(io.StringIO("".join(rows))):
        v = float(x["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[x["Metric Unit"]]
        a = agg.setdefault(x["Kernel Name"][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, (c, t) in agg.items():
        print("| `%s` | %d | %.1f | %.1f%% | %.2f |" % (k, c, t, 100 * t / tot, t / c))


if __name__ == "__main__":
    main(sys.argv[1])
