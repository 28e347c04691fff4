#!/usr/bin/env python3
"""debug helper: per-macroblock difference map between the GPU library and another backend for a .264 file"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from checkers import decode_bytes
import numpy as np
data = open(sys.argv[1], "rb").read(); other = sys.argv[2] if len(sys.argv) > 2 else "ref"
ref, _ = decode_bytes(data, other)
for rep in range(int(os.environ.get("REPS", "2"))):
    gpu, _ = decode_bytes(data, "gpu")
    bad = [i for i, (g, p) in enumerate(zip(gpu, ref)) if g[3] != p[3]]
    print("rep", rep, "bad frames", bad)
    for i in bad[:2]:
        g, p = gpu[i], ref[i]; w, h = g[1], g[2]
        a = np.frombuffer(g[3], np.uint8); b = np.frombuffer(p[3], np.uint8)
        Y = a[:w * h].reshape(h, w) != b[:w * h].reshape(h, w)
        C = (a[w * h:].reshape(2, h // 2, w // 2) != b[w * h:].reshape(2, h // 2, w // 2)).any(axis=0)
        print("frame", i, "id", g[0], "luma diffs", int(Y.sum()), "chroma diffs", int(C.sum()))
        ys, xs = np.nonzero(Y)
        if len(ys): print(" luma bbox x", xs.min(), xs.max(), "y", ys.min(), ys.max(), "first", (int(xs[0]), int(ys[0])), "gpu", a[:w*h].reshape(h, w)[ys[0], xs[0]], "ref", b[:w*h].reshape(h, w)[ys[0], xs[0]])
        ys, xs = np.nonzero(C)
        if len(ys): print(" chroma bbox x", xs.min(), xs.max(), "y", ys.min(), ys.max())
        mbs = sorted({(int(x) // 16, int(y) // 16) for y, x in zip(*np.nonzero(Y))})
        print(" luma MBs", mbs[:40])
