#!/usr/bin/env python
"""Where the time of ONE graph-launched substep goes: needs a -DBENDY_TIMESTAMPS build (BENDY2D_B200_LIB), in which the
four kernels of the critical path stamp %globaltimer per SM at entry / after the dependency wait / at their end.
Prints, relative to the first entry of the link kernel: first entry, first start of work, last end - medians over
several one-substep updates.  usage: substep_timeline.py [updates_before (3 = early, 28 = late)] [scene c3|c5strip]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import Solver, scenes
from bendy2d_b200 import _lib

L = _lib.lib()
before = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = scenes.c3_softbody_field()
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
sv = Solver()
sc.load_into(sv)
sv.update(sc.dt, n=before)
sv.synchronize()
sv.set_sub_steps(1)  # (an odd count: the update also pays the hand-back memset of the fill array, behind the narrowphase)
dt1 = float(np.float32(1 / 120.0))
sv.update(dt1, n=4)  # captures the one-substep graph
sv.synchronize()
names = ["links_local", "(scan)", "(scatter)", "narrowphase"]
rows = []
buf = (ctypes.c_ulonglong * (4 * 3 * 256))()
for rep in range(15):
    assert L.bendy_debug_ts_reset() == 0
    sv.update(dt1, n=1)
    sv.synchronize()
    assert L.bendy_debug_ts_read(buf) == 0
    a = np.frombuffer(buf, dtype=np.uint64).reshape(4, 3, 256).astype(np.float64)
    t0 = a[0, 0].min()
    row = []
    for k in range(4):
        used = a[k, 2] > 0
        if not used.any():  # kernel not in this build (the counting sort's scan / scatter after round 2's cell slots)
            row += [np.nan] * 5
            continue
        row += [a[k, 0][used].min() - t0, a[k, 1][used].min() - t0, np.median(a[k, 1][used]) - t0, a[k, 2][used].max() - t0,
                np.median(a[k, 2][used]) - t0]
    rows.append(row)
med = np.median(np.array(rows), axis=0) / 1000.0
print(f"C3 after {before} updates: one graph-launched substep, us after the first CTA of the link kernel entered")
print(f"{'kernel':12s} {'first entry':>11s} {'first work':>10s} {'median SM first work':>20s} {'last end':>9s} {'median SM last end':>18s}")
for k in range(4):
    e, s, sm, x, xm = med[5 * k:5 * k + 5]
    print(f"{names[k]:12s} {e:11.1f} {s:10.1f} {sm:20.1f} {x:9.1f} {xm:18.1f}")
