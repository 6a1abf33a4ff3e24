#!/usr/bin/env python
"""Where does the C3 substep time go?  Times the graph-mode substep of C3 variants (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import Solver, scenes


def run(name, sc, n_updates=25, sub=8):
    sc.sub_steps, sc.dt = sub, float(np.float32(sub / 120.0))
    sv = Solver()
    sc.load_into(sv)
    sv.update(sc.dt, n=3)
    sv.synchronize()
    sv.timer_start()
    sv.update(sc.dt, n=n_updates)
    ms = sv.timer_stop()
    us = ms * 1000 / (n_updates * sub)
    print(f"{name:42s} {sc.n_points:9d} pts  {us:8.1f} us/substep  {sc.n_points / us * 1e6:.3e} pss/s  {sv.schedule_info()['kernels_per_substep']} kernels")


if __name__ == "__main__":
    run("C3 full", scenes.c3_softbody_field())
    run("C3 no polygons", scenes.c3_softbody_field(50, 40, 200, 0))
    run("C3 no circles", scenes.c3_softbody_field(50, 40, 0, 500))
    run("C3 bodies only", scenes.c3_softbody_field(50, 40, 0, 0))
    sc = scenes.c3_softbody_field(50, 40, 0, 0)
    sc.particle_radius = 0.0
    run("C3 bodies only, no discs (links+K1)", sc)
    run("2M bodies only (100x40)", scenes.softbody_field(100, 40, (0.0, 0.0, 1024.0, 512.0), (56.0, 8.0), 0, 0, 1, "2M", 200))
    run("4M bodies only (100x80)", scenes.softbody_field(100, 80, (0.0, 0.0, 1024.0, 1024.0), (56.0, 8.0), 0, 0, 1, "4M", 200))
