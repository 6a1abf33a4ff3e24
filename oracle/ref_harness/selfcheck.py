#!/usr/bin/env python
"""Dry run of the harness pipeline WITHOUT Rust: a stand-in for src/main.rs that reads the same scene files, runs the
C++ oracle instead of the crate and writes the same output format; compare.py must then report OK three times.  It
proves the plumbing (scene export, update counts, dt, output parsing), not the pin itself: that needs cargo (run.sh)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import bo  # noqa: E402

import compare  # noqa: E402
import export_inputs  # noqa: E402


def fl(tok):
    return float(np.uint32(int(tok, 16)).view(np.float32))


def hx(v):
    return f"{np.float32(v).view(np.uint32):08x}"


def run_scene(src, dst):
    o = bo.OracleSolver()
    n_updates, dt = 0, 0.0
    for line in open(src):
        t = line.split()
        if t[0] == "bounds":
            o.set_bounds(*[fl(x) for x in t[1:5]])
        elif t[0] == "gravity":
            o.set_gravity(fl(t[1]), fl(t[2]))
        elif t[0] == "particle":
            o.add_particles(np.array([[fl(t[1]), fl(t[2])]], np.float32))
        elif t[0] == "plink":
            o.add_particle_links(np.array([[int(t[1]), int(t[2])]], np.uint32), np.array([fl(t[3])], np.float32))
        elif t[0] == "circle":
            o.add_circle(np.array([fl(t[1]), fl(t[2])], np.float32), fl(t[3]))
        elif t[0] == "clink":
            o.add_circle_link(int(t[1]), int(t[2]), fl(t[3]))
        elif t[0] == "polygon":
            n = int(t[2])
            pts = np.array([[fl(t[3 + 2 * i]), fl(t[4 + 2 * i])] for i in range(n)], np.float32)
            o.add_polygon_new(pts, t[1] == "1")
        elif t[0] == "run":
            n_updates, dt = int(t[1]), fl(t[2])
    for _ in range(n_updates):
        o.update(dt)
    with open(dst, "w") as f:
        if o.particle_len():
            p, q = o.particles()
            for a, b in zip(p, q):
                f.write(f"p {hx(a[0])} {hx(a[1])} {hx(b[0])} {hx(b[1])}\n")
        if o.circle_len():
            p, q, _ = o.circles()
            for a, b in zip(p, q):
                f.write(f"c {hx(a[0])} {hx(a[1])} {hx(b[0])} {hx(b[1])}\n")
        for k in range(o.polygon_len()):
            p, q, c = o.polygon(k)
            for a, b in zip(p, q):
                f.write(f"g {k} {hx(a[0])} {hx(a[1])} {hx(b[0])} {hx(b[1])}\n")
            f.write(f"gc {k} {hx(c[0])} {hx(c[1])}\n")


def main():
    with tempfile.TemporaryDirectory() as d:
        export_inputs.main(os.path.join(d, "inputs"))
        os.makedirs(os.path.join(d, "outputs"))
        for s in ("c1_reference_order", "circle_pile", "polygon_heap"):
            run_scene(os.path.join(d, "inputs", s + ".txt"), os.path.join(d, "outputs", s + ".txt"))
        compare.main(os.path.join(d, "outputs"))


if __name__ == "__main__":
    main()
