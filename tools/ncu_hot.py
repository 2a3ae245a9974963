#!/usr/bin/env python
"""Hot SASS lines of one kernel in an .ncu-rep (needs -lineinfo / --import-source on).
usage: python tools/ncu_hot.py file.ncu-rep kernel-regex rows [min_per_row]"""
import csv
import io
import re
import subprocess
import sys

rep, rx, nrows = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
parts = re.split(r'(?m)^"Kernel Name",', txt)
part = parts[1]
lines = part.split("\n")
print(lines[0][:100])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) >= len(hdr)]
st = sum(int(r[idx["Warp Stall Sampling (All Samples)"]] or 0) for r in body) or 1
tot = 0
for i, r in enumerate(body):
    n = int(r[idx["Instructions Executed"]])
    tot += n
    if n / nrows >= thr:
        print(f"{i:4d} {n / nrows:7.1f} {100 * int(r[idx['Warp Stall Sampling (All Samples)']] or 0) / st:5.1f}% "
              f"{r[idx['Avg. Threads Executed']]:>4} {r[idx['Source']][:90]}")
print("total per row", tot / nrows)
