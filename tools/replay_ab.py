#!/usr/bin/env python3
"""kernel-only replay timing for A/B experiments (env switches are read once per process)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
S = int(os.environ.get("S", "16")); steps = int(os.environ.get("STEPS", "3"))
os.environ["E264B_KEEP"] = "1"
bufs = bench.generate_streams([2000 + i for i in range(S)], 60, "/tmp/e264_bench")
lib = bench.BenchLib(os.path.join(bench.ROOT, "tools", "libe264bench.so"))
core = ctypes.CDLL(os.path.join(bench.ROOT, "edge264_b200", "libedge264_b200.so"))
core.e264b_of_decoder.restype = ctypes.c_void_p; core.e264b_of_decoder.argtypes = [ctypes.c_void_p]
core.e264b_replay.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint64)]
secs, frames, sums, decs = lib.run(bufs, min(S, 16), keep=True)
devs = (ctypes.c_void_p * S)(*[core.e264b_of_decoder(decs[i]) for i in range(S)])
ms = ctypes.c_float(); msr = ctypes.c_float(); nl = ctypes.c_uint64()
core.e264b_replay(devs, S, 1, ctypes.byref(ms), None, ctypes.byref(nl))
core.e264b_replay(devs, S, steps, ctypes.byref(ms), ctypes.byref(msr), ctypes.byref(nl))
n = sum(frames) * steps
print(f"{os.environ.get('TAG','')}: total {ms.value/steps:.1f} ms/step  recon-only {msr.value/steps:.1f} ms/step  -> {n/ms.value*1000:.0f} fps (e2e decode of warm-up: {sum(frames)/secs:.0f} fps)")
