#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, iv, ig, ib = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ik].split("(")[0].replace("void ", "")
    a = agg.setdefault(k, [0, 0.0, r[ig], r[ib]])
    a[0] += 1
    a[1] += float(r[iv]) / 1e6
tot = sum(a[1] for a in agg.values())
print("total %.2f ms over %d launches (per-launch times under ncu are serialised and cold-cache: shares, not absolutes)" % (tot, len(rows) - 1))
print("%-28s %5s %10s %10s %7s  %s" % ("kernel", "n", "total ms", "ms/launch", "share", "grid x block"))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-28s %5d %10.3f %10.4f %6.1f%%  %s x %s" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, a[2], a[3]))
