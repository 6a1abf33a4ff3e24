#!/usr/bin/env python
"""Discs per broadphase cell of C3 (h = 0.42 = 4.2 r) after N updates: how many slots does a cell need?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import Solver, scenes

sc = scenes.c3_softbody_field()
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
sv = Solver()
sc.load_into(sv)
done = 0
for upto in [int(a) for a in sys.argv[1:]] or [3, 15, 28, 60]:
    sv.update(sc.dt, n=upto - done)
    done = upto
    p, _ = sv.read_particles()
    for h in (0.42, 0.21):
        nx = int(np.ceil(512 / h))
        c = np.clip((p[:, 1] / h).astype(np.int64), 0, nx - 1) * nx + np.clip((p[:, 0] / h).astype(np.int64), 0, nx - 1)
        cnt = np.bincount(c, minlength=nx * nx)
        hist = np.bincount(cnt)
        over = {cap: int(np.maximum(cnt - cap, 0).sum()) for cap in (4, 6, 8, 12, 16)}
        print(f"after {upto} updates, h={h}: occupied cells {int((cnt > 0).sum())}, max {cnt.max()}, discs beyond cap {over}, "
              f"cells by count {hist[:20].tolist()}")
