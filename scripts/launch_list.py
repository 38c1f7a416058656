#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches.csv).
usage: python scripts/launch_list.py launches.csv > launch_list.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
ix = {h: i for i, h in enumerate(rows[0])}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0][:60]
    tot[name] += float(r[ix["Metric Value"]]) / 1e6
    cnt[name] += 1
probe = "k_dfma_peak"
total = sum(v for k, v in tot.items() if probe not in k)
print("%-60s %9s %14s %8s" % ("kernel", "launches", "total_ms", "share"))
for k, v in tot.most_common():
    print("%-60s %9d %14.3f %8s" % (k, cnt[k], v, "-" if probe in k else "%.2f%%" % (100 * v / total)))
