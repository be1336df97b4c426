#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            ns = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        key = (r[ik][:90], r[ig], r[ib])
        a = agg.setdefault(key, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += ns
        a[2] = min(a[2], ns)
        a[3] = max(a[3], ns)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | grid | block | launches | total ms | share | min us | max us |")
    print("|---|---|---|---|---|---|---|---|")
    for (k, g, b), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %s | %s | %d | %.3f | %.1f%% | %.1f | %.1f |" % (k, g, b, a[0], a[1] / 1e6, 100 * a[1] / tot, a[2] / 1e3, a[3] / 1e3))


if __name__ == "__main__":
    main()
