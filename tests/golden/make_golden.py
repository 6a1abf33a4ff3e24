#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference (Rust) cannot run in this environment, so these vectors are ORACLE outputs: they do
not pin the oracle to the reference (the KATs in tests/test_oracle_kat.py do that); they freeze the
oracle's behaviour on whole scenes so that an accidental change to it - or to the CUDA path, which
is compared against the same files on the GPU box - is caught.   Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bendy2d_b200 import scenes  # noqa: E402
from oracle import bo  # noqa: E402

f32 = np.float32


def c1_reference_order(n_updates=30):
    """C1 exactly as the reference runs it: insertion-order links, extensions off."""
    from helpers import oracle_from_scene

    sc = scenes.c1_softbody_blob()
    o = oracle_from_scene(sc)
    for _ in range(n_updates):
        o.update(sc.dt)
    pos, prev = o.particles()
    cp, cq, cr = o.circles()
    return dict(pos=pos, prev=prev, circle_pos=cp, circle_prev=cq, n_updates=n_updates)


def circle_pile(n_updates=40):
    rng = np.random.default_rng(3)
    o = bo.OracleSolver()
    o.set_bounds(0, 0, 30, 30)
    pos = rng.uniform(5, 25, size=(60, 2)).astype(f32)
    rad = rng.uniform(0.8, 2.5, size=60).astype(f32)
    for p, r in zip(pos, rad):
        o.add_circle(p, float(r))
    o.add_circle_link(0, 1, 4.0)
    o.add_circle_link(1, 5, 3.0)
    for _ in range(n_updates):
        o.update(1 / 120)
    cp, cq, _ = o.circles()
    return dict(init_pos=pos, radius=rad, pos=cp, prev=cq, n_updates=n_updates)


def polygon_heap(n_updates=160):
    rng = np.random.default_rng(11)
    o = bo.OracleSolver()
    o.set_bounds(0, 0, 40, 40)
    verts, statics = [], []
    for k in range(14):
        c = np.array([6.0 + 4.5 * (k % 6) + rng.uniform(-0.3, 0.3), 8.0 + 7.0 * (k // 6) + rng.uniform(-0.3, 0.3)])
        nv = int(rng.integers(3, 7))
        ang = rng.uniform(0, 6.28) + 2 * np.pi * np.arange(nv) / nv
        pts = (c + rng.uniform(1.6, 2.6) * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(f32)
        verts.append(pts)
        statics.append(k % 5 == 4)
        o.add_polygon_new(pts, statics[-1])
    for _ in range(n_updates):
        o.update(1 / 120)
    out = dict(n_updates=n_updates, n_poly=len(verts), statics=np.array(statics))
    for k, pts in enumerate(verts):
        p, q, c = o.polygon(k)
        out[f"init_{k}"], out[f"pos_{k}"], out[f"prev_{k}"], out[f"center_{k}"] = pts, p, q, c
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "c1_reference_order.npz"), **c1_reference_order())
    np.savez_compressed(os.path.join(HERE, "circle_pile.npz"), **circle_pile())
    np.savez_compressed(os.path.join(HERE, "polygon_heap.npz"), **polygon_heap())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
