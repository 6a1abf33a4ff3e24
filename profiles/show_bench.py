#!/usr/bin/env python
"""Pretty-print the bench.py JSON line(s) in a file."""
import json
import sys

for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("reference:", d["value"], d["cpu_baseline"]["sample"])
        continue
    r = d["roofline"]
    print(f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.4f} ({d['ms_per_step'] * 1000 / d['config']['substeps_per_step']:.1f} us/substep)"
          f"  warm-L2 {d['value_warm_l2']:.4g}  e2e {d['e2e']['value'] if d['e2e'] else None}  launches {d['gpu_launches']}")
    print(f"  substep alg bytes {r['substep']['alg_bytes'] / 1e6:.1f} MB  achieved {r['substep']['achieved']:.0f} GB/s  frac {r['substep']['frac']:.3f}"
          f"  | dominant {r['kernel']} {r['achieved']:.0f} GB/s frac {r['frac']:.3f}  clocks {d['clocks']}")
    for k, v in r["kernels"].items():
        print(f"    {k:14s} {v['ms_per_substep'] * 1000:7.1f} us/substep  x{v['launches_per_substep']:.0f}  share {v['share'] * 100:5.1f}%  {v.get('alg_GBps', 0):7.0f} GB/s")
    if d.get("cpu_baseline"):
        print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"])
