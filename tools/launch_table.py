#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: launch_table.py <csv> [first|second|all] [md]
The bench runs warm-up steps first, so `second` (default) takes the second half of the list = the timed step."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    part = sys.argv[2] if len(sys.argv) > 2 else "second"
    md = len(sys.argv) > 3 and sys.argv[3] == "md"
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = []
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "us" else v / 1e6 if r[ui] == "ns" else v
        data.append((r[ki], v))
    half = len(data) // 2
    data = data[half:] if part == "second" else data[:half] if part == "first" else data
    agg = collections.OrderedDict()
    for k, v in data:
        k = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:60]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    if md:
        print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print("| %s | %d | %.3f | %.1f%% |" % (k, a[0], a[1], 100 * a[1] / tot))
        else:
            print("%-62s %3d %8.3f ms %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
    if md:
        print("| **total** | %d | %.3f | 100%% |" % (len(data), tot))
    else:
        print("total %.3f ms, %d launches" % (tot, len(data)))


if __name__ == "__main__":
    main()
