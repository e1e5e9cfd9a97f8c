#!/usr/bin/env python
"""profiles/<tag>_ncu_full_<workload>.txt (scripts/ncu_summary.py output) -> profiles/traffic.json:
per workload, dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch (last captured launch),
which bench.py reports as roofline.traffic.  usage: python scripts/make_traffic_json.py r01"""
import glob
import json
import os
import re
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
out = {}
for path in sorted(glob.glob(os.path.join(root, f"{tag}_ncu_full_*.txt"))):
    w = os.path.basename(path)[len(tag) + len("_ncu_full_"):-4]
    rd = wr = name = dur = None
    for ln in open(path):
        m = re.match(r"(\S+)\s+(.*?)\s+(\S+)\s*$", ln.rstrip("\n"))
        if ln.startswith("Kernel Name"):
            name = ln[len("Kernel Name"):].strip()
        elif ln.startswith("dram__bytes_read.sum"):
            v, u = ln.split()[1:3]
            rd = float(v) * UNITS[u]
        elif ln.startswith("dram__bytes_write.sum"):
            v, u = ln.split()[1:3]
            wr = float(v) * UNITS[u]
        elif ln.startswith("gpu__time_duration.sum"):
            dur = " ".join(ln.split()[1:3])
    if rd is not None and wr is not None:
        out[w] = {"bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": name, "ncu_duration": dur,
                  "source": f"profiles/{os.path.basename(path)}"}
json.dump(out, open(os.path.join(root, "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
