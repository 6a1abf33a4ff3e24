#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from
`ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --print-source cuda,sass > f.csv`.
usage: hot_lines.py f.csv [top=40]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
lines = {}
first_kernel = True
n_hdr = 0
for r in rows:
    if r and r[0] == "Line No":
        n_hdr += 1
        if n_hdr > 1:
            break  # only the first captured instance
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r or r[0] in ("File Path", "Function Name"):
        continue
    if r[0] != "":  # a source line row
        try:
            ln = int(r[0])
        except ValueError:
            continue
        num = lambda v: int(v) if v not in ("", "-") else 0
        inst = num(r[hdr["Instructions Executed"]])
        tinst = num(r[hdr["Thread Instructions Executed"]])
        samp = num(r[hdr["# Samples"]])
        e = lines.setdefault(ln, [r[1], 0, 0, 0])
        e[1] += inst
        e[2] += tinst
        e[3] += samp
tot = sum(v[1] for v in lines.values())
tsamp = sum(v[3] for v in lines.values())
print(f"total warp instructions {tot}, samples {tsamp}")
for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][3])[:top]:
    print(f"{ln:5d} inst {v[1]:9d} ({100.0 * v[1] / max(tot, 1):4.1f}%) lanes {v[2] / max(v[1], 1):4.1f} samples {v[3]:6d} ({100.0 * v[3] / max(tsamp, 1):4.1f}%)  {v[0].strip()[:110]}")
