#!/usr/bin/env python
"""List the SASS of one kernel from `ncu --page source --csv` with executed-instruction counts and
stall samples; usage: sass_hot.py file.csv [min_share_pct]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > max(ia, ie, isamp) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in body)
tots = sum(int(r[isamp]) for r in body)
print("total warp-inst", tot, "samples", tots, "sass lines", len(body))
for i, r in enumerate(body):
    e, sm = int(r[ie]), int(r[isamp])
    print(f"{i:4d} {e / 1000:8.0f}k {100.0 * e / tot:5.1f}% {sm:5d} {100.0 * sm / max(tots, 1):5.1f}%  {r[ia].strip()[:100]}")
