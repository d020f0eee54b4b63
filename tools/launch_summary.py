#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`).
Usage: python tools/launch_summary.py X.csv [top_n]"""
import collections
import csv
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v, u = float(r[vi].replace(",", "")), r[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(r[ki][:72], [0, 0.0])
        a[0] += 1
        a[1] += v
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-74s n=%4d total %9.1f us  avg %8.1f us" % (k, n, t, t / n))


if __name__ == "__main__":
    main()
