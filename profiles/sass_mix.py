#!/usr/bin/env python
"""Instruction mix of the hot kernels from `cuobjdump -sass` of the built library (no GPU needed).
usage: sass_mix.py [lib] > profiles/r2_sass_hot_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bendy2d_b200", "lib", "libbendy2d_b200.so")
HOT = ["k2_narrow_contact_integrateILb0ELb1ELb1", "k2_scatterILb0", "k3_links_localILb0ELb1ELi0", "k_polygons_fused", "k2_scanE",
       "k_circles_exact"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print(f"cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a), instruction mix of the hot kernels.")
print("No tensor-core (MMA) and no TMA (UBLKCP/UTMALDG) instructions: the path is gather/scatter f32 work.\n")
cur, mix = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if any(h in m.group(1) for h in HOT) else None
        if cur:
            mix[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if cur and m:
        mix[cur][m.group(1)] += 1
for name, c in mix.items():
    n = sum(c.values())
    g = lambda *ops: sum(v for k, v in c.items() if k in ops)
    print(name)
    print(f"  {n} SASS instructions; FFMA {c['FFMA']} (contraction is off: the FFMAs are the div.rn / sqrt.rn refinement sequences "
          f"and the shared-reciprocal normalize), MUFU {c['MUFU']}, LDG {c['LDG']}, STG {c['STG']}, LDS {c['LDS']}, "
          f"ATOM/RED {g('ATOM', 'ATOMG', 'ATOMS', 'RED', 'REDG')}, BAR {c['BAR']}, "
          f"HMMA/UTC*MMA {sum(v for k, v in c.items() if 'MMA' in k)}")
    print("  mix: " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)) + "\n")
