#!/usr/bin/env python
"""Graph-mode substep time of C3 early (free fall) and late (piled up) for the library selected by
BENDY2D_B200_LIB / BENDY_* env knobs.  usage: quick_c3.py [label] [pack_points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import Solver, scenes

label = sys.argv[1] if len(sys.argv) > 1 else "default"
sc = scenes.c3_softbody_field()
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
sv = Solver()
sc.load_into(sv)
if len(sys.argv) > 2:
    sv.set_plan_params(int(sys.argv[2]), 0)
if os.environ.get("QUICK_C3_CELL"):  # broadphase cell size (default: auto = 4.2 r_p)
    sv.set_grid_cell(float(os.environ["QUICK_C3_CELL"]))
sv.update(sc.dt, n=3)
sv.synchronize()
out = []
for n in (12, 10, 3):  # updates 3..15 (early), 15..25 (transition), 25..28 (late)
    sv.timer_start()
    sv.update(sc.dt, n=n)
    out.append(sv.timer_stop() * 1000 / (n * 8))
total = (out[0] * 12 + out[1] * 10 + out[2] * 3) / 25
print(f"{label:28s} early {out[0]:6.1f}  mid {out[1]:6.1f}  late {out[2]:6.1f}  avg(25 steps) {total:6.1f} us/substep")
