#!/usr/bin/env python
"""profiles/traffic.json from an ncu `--set full` capture: DRAM bytes (read + write) per launch of the
dominant kernel, keyed the way bench.py looks it up (<workload>_<dtype>_<kernel key>), stamped with the
capture file and the commit it was taken at.
usage: python tools/make_traffic.py <file.ncu-rep> <kernel-regex> <key> [commit]"""
import csv
import json
import os
import re
import subprocess
import sys

rep, rx, key = sys.argv[1], re.compile(sys.argv[2]), sys.argv[3]
commit = sys.argv[4] if len(sys.argv) > 4 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


vals = []
for r in data:
    if rx.search(r[idx["Kernel Name"]]):
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        vals.append((rd + wr, float(r[idx["gpu__time_duration.sum"]].replace(",", "")), units[idx["gpu__time_duration.sum"]]))
if not vals:
    raise SystemExit("no kernel matches")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
b = sum(v[0] for v in vals) / len(vals)
tj[key] = {"bytes": int(b), "launches_averaged": len(vals), "kernel_time": f"{vals[0][1]} {vals[0][2]} (under ncu)",
           "source": f"{os.path.basename(rep)}, ncu --set full --clock-control none, commit {commit}"}
json.dump(tj, open(path, "w"), indent=1, sort_keys=True)
print(key, tj[key])
