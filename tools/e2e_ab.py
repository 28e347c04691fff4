#!/usr/bin/env python3
"""e2e decode throughput (host-bound part of bench.py) for several application-thread counts, with the process CPU time
per frame (user / sys) so that host overheads beyond bitstream parsing show up.  GPU box only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cfg = bench.CONFIGS["1080p"]
lib = bench.BenchLib(os.path.join(bench.ROOT, "tools", "libe264bench.so"))
ref = bench.BenchLib(os.path.join(bench.ROOT, "oracle", "_ref", "libe264bench_ref.so"))
bufs = bench.generate_streams(cfg, [2000 + i for i in range(32)], 60, "/tmp/e264_bench")
def cpu(): t = os.times(); return t.user, t.system
def run(L, name, b, T, env=None):
    for k, v in (env or {}).items(): os.environ[k] = v
    L.run(b, T)
    u0, s0 = cpu(); t0 = time.time()
    n = 0
    for _ in range(2):
        s, fr, _, d = L.run(b, T); n += sum(fr)
    u1, s1 = cpu(); w = time.time() - t0
    print(f"{name} streams {len(b)} threads {T} {env or ''}: {n / w:.0f} fps; CPU per frame: user {1000 * (u1 - u0) / n:.2f} ms sys {1000 * (s1 - s0) / n:.2f} ms; CPUs busy {(u1 - u0 + s1 - s0) / w:.1f}", flush=True)
for T in (32, 16, 20, 24, 48):
    run(lib, "gpu", bufs, T)
run(lib, "gpu", bufs, 16, {"E264_BENCH_DEC_THREADS": "1"})
run(lib, "gpu", bufs[:16], 16, {"E264_BENCH_DEC_THREADS": "1"})
os.environ["E264_BENCH_DEC_THREADS"] = "0"
for T in (16, 32):
    run(ref, "ref", bufs, T)
