#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list (cold-cache, serialised: compare SHARES).
   python scripts/launch_summary.py gpurun_out/launches_X.csv [steps] > profiles/X_launches_summary.txt"""
import collections
import csv
import sys

steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    n += 1
    name = row["Kernel Name"].split("(")[0][:100]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if "at::" not in k and "cutlass" not in k and "cub::" not in k)
print("%d launches over %d steps (1 warm-up + %d), serialised total %.1f us = %.2f ms/step; this library's kernels %.1f%% of it"
      % (n, steps, steps - 1, tot, tot / steps / 1e3, 100 * ours / tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%10.1f us %5d launches %5.1f%%  %s" % (a[1], a[0], 100 * a[1] / tot, k))
