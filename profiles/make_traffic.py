#!/usr/bin/env python
"""profiles/traffic.json from `ncu -i prof.ncu-rep --page raw --csv`: DRAM bytes (read + write) per launch of each kernel
class, keyed by the WORKLOAD the capture ran (bench.py only quotes a figure for the workload it was captured on, N = 1).
usage: make_traffic.py raw.csv workload source-note"""
import csv
import json
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
workload, note = sys.argv[2], sys.argv[3]
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[idx[name]].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[name]], 1)


names = {"k2_narrow_contact_integrate": "narrowphase", "k3_links_local": "links_local", "k2_scatter": "grid_build_scatter",
         "k2_scan": "grid_build_scan", "k_polygons_fused": "poly_prep", "k_circles_exact": "circle_pass", "k2_circle_bin": "circles",
         "k1_integrate_range": "integrate"}
out = {}
for r in rows[2:]:
    kn = r[idx["Kernel Name"]]
    for key, cls in names.items():
        if key in kn and cls not in out:
            out[cls] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
out["grid_build"] = out.get("grid_build_scatter", 0) + out.get("grid_build_scan", 0)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
allw = json.load(open(dst)) if os.path.exists(dst) else {}
if "source" in allw and not isinstance(allw.get("source"), dict):
    allw = {}  # round-1 layout
allw[workload] = dict(out, source=note)
json.dump(allw, open(dst, "w"), indent=1)
print(json.dumps(allw[workload], indent=1))
