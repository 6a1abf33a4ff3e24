#!/usr/bin/env python
"""Writes the scene files of the three golden vectors (tests/golden/*.npz) for the Rust harness: the same initial
states, link lists and update counts tests/golden/make_golden.py feeds the oracle.  usage: export_inputs.py outdir"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bendy2d_b200 import scenes  # noqa: E402

f32 = np.float32


def hx(v):
    return f"{np.float32(v).view(np.uint32):08x}"


def write(path, bounds, particles=(), plinks=(), circles=(), clinks=(), polygons=(), n_updates=0, dt=0.0, gravity=(0.0, 98.2)):
    with open(path, "w") as f:
        f.write("bounds " + " ".join(hx(v) for v in bounds) + "\n")
        f.write("gravity " + " ".join(hx(v) for v in gravity) + "\n")
        for p in particles:
            f.write(f"particle {hx(p[0])} {hx(p[1])}\n")
        for a, b, ln in plinks:
            f.write(f"plink {int(a)} {int(b)} {hx(ln)}\n")
        for x, y, r in circles:
            f.write(f"circle {hx(x)} {hx(y)} {hx(r)}\n")
        for a, b, ln in clinks:
            f.write(f"clink {int(a)} {int(b)} {hx(ln)}\n")
        for pts, st in polygons:
            f.write(f"polygon {1 if st else 0} {len(pts)} " + " ".join(f"{hx(x)} {hx(y)}" for x, y in pts) + "\n")
        f.write(f"run {n_updates} {hx(dt)}\n")


def main(out):
    os.makedirs(out, exist_ok=True)
    G = os.path.join(ROOT, "tests", "golden")
    # C1 exactly as the reference runs it: 8 substeps of (1/60)/8 = 8 update() calls of dt/8 per frame is arithmetically
    # the same as the oracle's sub_steps = 8 (x0.125 is exact); the golden was made with the scene's dt and sub_steps
    sc = scenes.c1_softbody_blob()
    g = np.load(os.path.join(G, "c1_reference_order.npz"))
    sub = sc.sub_steps
    write(os.path.join(out, "c1_reference_order.txt"), sc.bounds, sc.particles,
          [(a, b, ln) for (a, b), ln in zip(sc.links_ab, sc.links_len)],
          [(p[0], p[1], r) for p, r in zip(sc.circles_pos, sc.circles_r)],
          n_updates=int(g["n_updates"]) * sub, dt=np.float32(np.float32(sc.dt) * np.float32(1.0 / sub)))
    g = np.load(os.path.join(G, "circle_pile.npz"))
    write(os.path.join(out, "circle_pile.txt"), (0, 0, 30, 30),
          circles=[(p[0], p[1], r) for p, r in zip(g["init_pos"], g["radius"])],
          clinks=[(0, 1, 4.0), (1, 5, 3.0)], n_updates=int(g["n_updates"]), dt=np.float32(1 / 120))
    g = np.load(os.path.join(G, "polygon_heap.npz"))
    write(os.path.join(out, "polygon_heap.txt"), (0, 0, 40, 40),
          polygons=[(g[f"init_{k}"], bool(g["statics"][k])) for k in range(int(g["n_poly"]))],
          n_updates=int(g["n_updates"]), dt=np.float32(1 / 120))
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "inputs"))
