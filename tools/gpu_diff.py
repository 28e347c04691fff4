#!/usr/bin/env python3
"""debug helper: first differing sample between the GPU library and the oracle port on a generated stream"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from checkers import decode_bytes
import numpy as np
W, H = int(sys.argv[1]), int(sys.argv[2]); args = sys.argv[3:]
path = "/tmp/gpudiff.264"
subprocess.run(["tools/gen264", "-o", path, "-W", str(W), "-H", str(H)] + args, check=True, stderr=subprocess.DEVNULL)
data = open(path, "rb").read()
port, _ = decode_bytes(data, "port")
for rep in range(3):
    gpu, _ = decode_bytes(data, "gpu")
    msg = "OK"
    for i, (g, p) in enumerate(zip(gpu, port)):
        if g[3] != p[3]:
            a = np.frombuffer(g[3], np.uint8); b = np.frombuffer(p[3], np.uint8)
            w, h = g[1], g[2]; d = np.nonzero(a != b)[0]; o = int(d[0])
            if o < w * h: msg = f"frame {i} id {g[0]} Y y={o // w} x={o % w} mb=({(o % w) // 16},{(o // w) // 16}) gpu={a[o]} port={b[o]} ndiff={len(d)}"
            else:
                o -= w * h; pl = o // (w * h // 4); o %= w * h // 4
                msg = f"frame {i} id {g[0]} C{pl} y={o // (w // 2)} x={o % (w // 2)} mb=({(o % (w // 2)) // 8},{(o // (w // 2)) // 8}) gpu={a[o + w*h + pl*(w*h//4)]} ndiff={len(d)}"
            if os.environ.get("MAP"):
                Y1 = a[:w*h].reshape(h, w) != b[:w*h].reshape(h, w)
                for my in range(h // 16): print("".join("X" if Y1[my*16:my*16+16, mx*16:mx*16+16].any() else "." for mx in range(w // 16)))
            break
    print(" ".join(args), "| rep", rep, msg)
