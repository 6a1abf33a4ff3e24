#!/usr/bin/env python
"""profiles/traffic.json from `ncu -i prof.ncu-rep --page raw --csv`: DRAM bytes per launch of the main-chain
kernels (bench.py copies the dominant kernel's figure into roofline.traffic).  usage: make_traffic.py raw.csv"""
import csv
import json
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[idx[name]].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[name]], 1)


names = {"k2_narrow_contact_integrate": "narrowphase", "k3_links_local": "links_local", "k2_scatter": "grid_build_scatter",
         "k2_scan_fused": "grid_build_scan"}
out = {}
for r in rows[2:]:
    kn = r[idx["Kernel Name"]]
    for key, cls in names.items():
        if key in kn and cls not in out:
            out[cls] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum")}
            out[cls]["dram_bytes_total"] = out[cls]["dram_bytes_read"] + out[cls]["dram_bytes_write"]
traffic = {k: v["dram_bytes_total"] for k, v in out.items()}
traffic["grid_build"] = traffic.get("grid_build_scatter", 0) + traffic.get("grid_build_scan", 0)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
json.dump({"source": "ncu --set full (cache-control all: cold L2), per launch; profiles/r1_ncu_full_top_kernels.txt",
           **traffic, "detail": out}, open(dst, "w"), indent=1)
print(json.dumps(traffic, indent=1))
