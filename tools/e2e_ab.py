#!/usr/bin/env python3
"""e2e decode throughput for several stream/thread counts (host-bound part of bench.py)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
lib = bench.BenchLib(os.path.join(bench.ROOT, "tools", "libe264bench.so"))
ref = bench.BenchLib(os.path.join(bench.ROOT, "oracle", "_ref", "libe264bench_ref.so"))
bufs = bench.generate_streams([2000 + i for i in range(48)], 60, "/tmp/e264_bench")
for S, T in ((16, 16), (24, 24), (32, 32), (48, 48), (32, 16), (48, 24)):
    b = bufs[:S]
    lib.run(b, T)
    s, fr, _, d = lib.run(b, T); 
    s2, fr2, _, d2 = lib.run(b, T)
    print(f"gpu  streams {S} threads {T}: {sum(fr)/s:.0f} / {sum(fr2)/s2:.0f} fps", flush=True)
for S, T in ((16, 16), (32, 16), (32, 32)):
    b = bufs[:S]
    s, fr, _, _ = ref.run(b, T); s2, fr2, _, _ = ref.run(b, T)
    print(f"ref  streams {S} threads {T}: {sum(fr)/s:.0f} / {sum(fr2)/s2:.0f} fps", flush=True)
