#!/usr/bin/env python
"""profiles/<tag>_ncu_full_<name>.txt (scripts/ncu_summary.py output) and profiles/<tag>_traffic_c2_range.csv ->
profiles/traffic.json: per workload, dram__bytes_read.sum + dram__bytes_write.sum of ONE WHOLE STEP, which bench.py
reports as roofline.traffic.

  * scoring workloads: the step is emei_sumsq + the row kernel, each > L2, one launch each (last captured launch);
  * c2: the step's 22 MB of stores are still in the 126 MB L2 when ONE profiled launch ends, so a single-launch capture
    shows reads only.  The steady state is measured over a RANGE of 16 consecutive launches of the rotated ring
    (ncu --cache-control none: nothing is flushed between them, so the write-back of earlier launches is counted).
usage: python scripts/make_traffic_json.py r02"""
import csv
import json
import os
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")


def last_launch(name):
    path = os.path.join(root, f"{tag}_ncu_full_{name}.txt")
    if not os.path.exists(path):
        path = os.path.join(root, f"r01_ncu_full_{name}.txt")
        if not os.path.exists(path):
            return None
    rd = wr = kern = dur = None
    for ln in open(path):
        if ln.startswith("Kernel Name"):
            kern = ln[len("Kernel Name"):].strip()
        elif ln.startswith("dram__bytes_read.sum"):
            v, u = ln.split()[1:3]
            rd = float(v) * UNITS[u]
        elif ln.startswith("dram__bytes_write.sum"):
            v, u = ln.split()[1:3]
            wr = float(v) * UNITS[u]
        elif ln.startswith("gpu__time_duration.sum"):
            dur = " ".join(ln.split()[1:3])
    if rd is None:
        return None
    return {"read": rd, "write": wr, "kernel": kern, "ncu_duration": dur, "source": f"profiles/{os.path.basename(path)}"}


out = {}
# ---- scoring: sumsq + row kernel
ss = last_launch("sumsq")  # captured on 2^24 x 3 float32 actions (Hopper); HalfCheetah's pass reads twice as many
for w, rows, act_scale in (("c3_hopper", "c3_hopper", 1.0), ("c3_halfcheetah", "c3_halfcheetah", 2.0),
                           ("c3_hopper_seq", "c3_hopper_seq", 1.0), ("c3_halfcheetah_seq", "c3_halfcheetah_seq", 2.0)):
    k = last_launch(rows)
    if k is None:
        continue
    s_bytes = (ss["read"] + ss["write"]) * act_scale if ss else 0.0
    out[w] = {"bytes_per_launch": k["read"] + k["write"] + s_bytes, "row_kernel": k, "sumsq_bytes": s_bytes,
              "kernel": k["kernel"], "source": k["source"],
              "note": f"{k['source']} (row kernel: dram__bytes_read.sum + dram__bytes_write.sum) + the sum-of-squares pass "
                      f"({'profiles/' + tag + '_ncu_full_sumsq.txt' if ss else 'not captured'}"
                      + (", scaled x2 for 6-d actions" if act_scale != 1.0 else "") + "): one whole step, ncu --set full"}
for w in ("c1", "c4", "rollout", "rollout_rec", "c2_f64"):  # (c4_rollout was captured at 2^24 envs, not the bench size)
    k = last_launch(w)
    if k is not None:
        out[w] = {"bytes_per_launch": k["read"] + k["write"], "kernel": k["kernel"], "source": k["source"], "ncu_duration": k["ncu_duration"],
                  "note": f"{k['source']}: dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full; writes of a launch smaller than L2 may still sit in L2)"}
# ---- c2: range of consecutive launches
path = os.path.join(root, f"{tag}_traffic_c2_range.csv")
if os.path.exists(path):
    rows = [r for r in csv.reader(ln for ln in open(path) if ln.startswith('"'))]
    hdr = rows[0]
    iid, iname, ival, imet = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    per = {}
    for r in rows[1:]:
        per.setdefault(r[iid], {})[r[imet]] = float(r[ival].replace(",", ""))
    n = len(per)
    rd = sum(v["dram__bytes_read.sum"] for v in per.values()) / n
    wr = sum(v["dram__bytes_write.sum"] for v in per.values()) / n
    out["c2"] = {"bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "launches": n, "kernel": rows[1][iname],
                 "source": f"profiles/{os.path.basename(path)}",
                 "note": f"profiles/{os.path.basename(path)}: mean of dram__bytes_read.sum + dram__bytes_write.sum over {n} CONSECUTIVE launches of the "
                         "rotated 8-batch ring (ncu --cache-control none --clock-control none: caches are not flushed between launches, so the "
                         "write-back of earlier launches' stores is counted; a single profiled launch leaves its 22 MB of stores in L2)"}
json.dump(out, open(os.path.join(root, "traffic.json"), "w"), indent=1)
print(json.dumps({k: (v["bytes_per_launch"], v["note"][:60]) for k, v in out.items()}, indent=1))
