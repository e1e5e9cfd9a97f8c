#!/usr/bin/env python
"""Per-kernel share of a bench step from an ncu launch list (scripts/gpu_launchlist.sh):
usage: python scripts/launch_share.py profiles/r01_launches_<workload>.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
dur = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    if r[-3] == "gpu__time_duration.sum":
        name = r[4].split("(")[0]
        dur[name] += float(r[-1].replace(",", ""))
        cnt[name] += 1
tot = sum(dur.values())
for k, v in sorted(dur.items(), key=lambda kv: -kv[1]):
    print(f"{100 * v / tot:6.2f} %  {cnt[k]:4d} launches  {v / cnt[k] / 1e3:10.2f} us each  {k}")
