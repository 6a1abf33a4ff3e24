#!/usr/bin/env python
"""Compares the harness output (the REAL reference crate) with the oracle-frozen golden vectors bit for bit.
usage: compare.py outputs_dir      exit code 0 = the oracle is pinned by the reference on these scenes"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "tests", "golden")


def load(path):
    P, C, GP, GC = [], [], {}, {}
    for line in open(path):
        t = line.split()
        v = [np.uint32(int(x, 16)) for x in t[1:] if len(x) == 8]
        if t[0] == "p":
            P.append(v)
        elif t[0] == "c":
            C.append(v)
        elif t[0] == "g":
            GP.setdefault(int(t[1]), []).append([np.uint32(int(x, 16)) for x in t[2:]])
        elif t[0] == "gc":
            GC[int(t[1])] = [np.uint32(int(x, 16)) for x in t[2:]]
    return np.array(P, np.uint32), np.array(C, np.uint32), GP, GC


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main(out):
    bad = 0
    P, C, _, _ = load(os.path.join(out, "c1_reference_order.txt"))
    g = np.load(os.path.join(G, "c1_reference_order.npz"))
    ok = np.array_equal(P[:, :2], bits(g["pos"])) and np.array_equal(P[:, 2:], bits(g["prev"])) and np.array_equal(C[:, :2], bits(g["circle_pos"]))
    print("c1_reference_order", "OK" if ok else "DIFFERS"); bad += not ok
    _, C, _, _ = load(os.path.join(out, "circle_pile.txt"))
    g = np.load(os.path.join(G, "circle_pile.npz"))
    ok = np.array_equal(C[:, :2], bits(g["pos"])) and np.array_equal(C[:, 2:], bits(g["prev"]))
    print("circle_pile", "OK" if ok else "DIFFERS"); bad += not ok
    _, _, GP, GC = load(os.path.join(out, "polygon_heap.txt"))
    g = np.load(os.path.join(G, "polygon_heap.npz"))
    ok = True
    for k in range(int(g["n_poly"])):
        a = np.array(GP[k], np.uint32)
        ok = ok and np.array_equal(a[:, :2], bits(g[f"pos_{k}"])) and np.array_equal(a[:, 2:], bits(g[f"prev_{k}"]))
        ok = ok and np.array_equal(np.array(GC[k], np.uint32), bits(g[f"center_{k}"]).ravel())
    print("polygon_heap", "OK" if ok else "DIFFERS"); bad += not ok
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main(sys.argv[1])
