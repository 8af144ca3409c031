#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, count / mean / last duration (us)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    else:
        sys.exit("no header")
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) > vi:
            agg.setdefault(r[ki][:72], []).append(float(r[vi].replace(",", "")))
    for n, v in agg.items():
        print(f"{n:72s} n={len(v):3d} mean={sum(v) / len(v) / 1e3:9.1f} us  last={v[-1] / 1e3:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
