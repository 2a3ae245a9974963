#!/usr/bin/env python
"""Stall samples of one kernel in an .ncu-rep, split into the SASS segments between BAR.SYNCs
(address order), with the dominant stall reasons per segment.
usage: python tools/ncu_phases.py file.ncu-rep kernel-regex"""
import csv
import io
import re
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
parts = [p for p in re.split(r'(?m)^"Kernel Name",', txt)[1:] if re.search(rx, p.split("\n")[0])]
part = parts[-1]   # last matching launch
lines = part.split("\n")
print(lines[0][:90])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) >= len(hdr)]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
S = "Warp Stall Sampling (All Samples)"
total = sum(int(r[idx[S]] or 0) for r in body) or 1
seg, segs = [], []
for r in body:
    seg.append(r)
    if "BAR.SYNC" in r[idx["Source"]]:
        segs.append(seg)
        seg = []
segs.append(seg)
for k, sg in enumerate(segs):
    st = sum(int(r[idx[S]] or 0) for r in sg)
    ins = sum(int(r[idx["Instructions Executed"]] or 0) for r in sg)
    wf = sum(int(r[idx["L1 Wavefronts Shared"]] or 0) for r in sg)
    rs = sorted(((sum(int(r[idx[h]] or 0) for r in sg), h) for h in reasons), reverse=True)[:4]
    top = max(sg, key=lambda r: int(r[idx[S]] or 0)) if sg else None
    print(f"seg {k:2d} n={len(sg):4d} stall {100 * st / total:5.1f}% inst {ins:>10d} smem_wf {wf:>10d}  "
          + " ".join(f"{h[6:]}={100 * v / total:.1f}" for v, h in rs if v)
          + (f"   | hot: {top[idx['Source']].strip()[:50]}" if top else ""))
