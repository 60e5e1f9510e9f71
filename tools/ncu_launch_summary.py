#!/usr/bin/env python3
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, rows = rows[hdr_i], rows[hdr_i + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(r[ui], 1.0)
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v * scale
    total = sum(v[1] for v in agg.values())
    print(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / v[0] / 1e3:10.1f} {100 * v[1] / total:6.1f}%")


if __name__ == "__main__":
    main()
