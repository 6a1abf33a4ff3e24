#!/usr/bin/env python
"""Eager (one launch per kernel) substeps of the C3 scene for an ncu capture of the particle chain.
usage: ncu ... python profiles/ncu_c3_chain.py [updates_before=0] [updates=1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import Solver, scenes

before = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = scenes.c3_softbody_field()
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
sv = Solver()
sc.load_into(sv)
sv.set_profiling(True)  # eager launches: every kernel is its own launch
if before:
    sv.update(sc.dt, n=before)
sv.synchronize()
sv.update(sc.dt, n=n)
sv.synchronize()
print("done", sv.stats())
