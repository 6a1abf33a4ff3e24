#!/usr/bin/env python
"""Per-kernel CUDA-event times of the C3 scene at chosen points of its evolution (GPU box).
Early the bodies are in free fall; late they are piled on circles, polygons and each other."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import Solver, scenes

sc = scenes.c3_softbody_field()
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
sv = Solver()
sc.load_into(sv)
if len(sys.argv) > 1:
    sv.set_grid_cell(float(sys.argv[1]))
print("grid", sv.grid())
done = 0
for upto in (3, 15, 21, 28):
    sv.update(sc.dt, n=upto - done)
    done = upto
    sv.synchronize()
    sv.timer_start()
    sv.update(sc.dt)
    graph_ms = sv.timer_stop()
    done += 1
    sv.set_profiling(True)
    sv.kernel_times(reset=True)
    sv.update(sc.dt)
    done += 1
    kt = sv.kernel_times(reset=True)
    sv.set_profiling(False)
    print("stats", sv.stats())
    print(f"after {upto} updates: graph {graph_ms * 1000 / 8:.1f} us/substep; eager per substep: " +
          ", ".join(f"{k} {v['ms'] * 1000 / 8:.1f}us x{v['launches'] // 8}" for k, v in kt.items() if v["launches"]))
