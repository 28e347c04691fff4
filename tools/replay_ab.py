#!/usr/bin/env python3
"""kernel-only replay timing for A/B experiments (env switches are read once per process)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
S = int(os.environ.get("S", "16")); steps = int(os.environ.get("STEPS", "3"))
os.environ["E264B_KEEP"] = "1"
bufs = bench.generate_streams(bench.CONFIGS["1080p"], [2000 + i for i in range(S)], 60, "/tmp/e264_bench")
lib = bench.BenchLib(os.path.join(bench.ROOT, "tools", "libe264bench.so"))
core = ctypes.CDLL(os.path.join(os.environ.get("E264_LIB_DIR", os.path.join(bench.ROOT, "edge264_b200")), "libedge264_b200.so"))   # variants: set LD_LIBRARY_PATH to the same directory
core.e264b_of_decoder.restype = ctypes.c_void_p; core.e264b_of_decoder.argtypes = [ctypes.c_void_p]
core.e264b_replay.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(bench.ReplayStats)]
secs, frames, sums, decs = lib.run(bufs, min(S, 16), keep=True)
devs = (ctypes.c_void_p * S)(*[core.e264b_of_decoder(decs[i]) for i in range(S)])
st = bench.ReplayStats(); T = int(os.environ.get("LAUNCH_THREADS", "8"))
core.e264b_replay(devs, S, 1, T, ctypes.byref(st))
core.e264b_replay(devs, S, steps, T, ctypes.byref(st))
n = sum(frames) * steps
per = " ".join(f"{n}={1000 * st.kernel_ms[k] / max(1, st.kernel_launches[k]):.0f}us" for k, n in ((1, "inter"), (2, "intra"), (3, "deblock")))
print(f"{os.environ.get('TAG','')}: total {st.ms_total/steps:.1f} ms/step ({"graph" if st.threads == 0 else str(st.threads) + " launch threads"}) -> {n/st.ms_total*1000:.0f} fps; mean span per launch: {per} (e2e decode of warm-up: {sum(frames)/secs:.0f} fps)")
