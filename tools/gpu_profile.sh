#!/bin/bash
# One GPU-box session: host facts, ncu launch list + full capture of both kernels, thread scaling of both arms.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-r1}
{ echo "nproc $(nproc)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | grep -E "Model name|Socket|Core|Thread" ; nvidia-smi -L; } > gpurun_out/host_$TAG.txt 2>&1
tools/gen264 -o /tmp/c2.264 -W 120 -H 68 -n 20 -s 2000 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv tools/b200_decode /tmp/c2.264 -q > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -s 12 -c 8 -o gpurun_out/prof_$TAG -f tools/b200_decode /tmp/c2.264 -q >> gpurun_out/ncu_$TAG.log 2>&1
if [ -n "$SCALING" ]; then
python - <<'PY' > gpurun_out/scaling_$1.txt 2>&1
import ctypes, os, sys, time
sys.path.insert(0, '.')
import bench
bufs = bench.generate_streams(bench.CONFIGS['1080p'], [2000 + i for i in range(16)], 60, '/tmp/e264_bench')
for name, path in (("reference", "oracle/_ref/libe264bench_ref.so"), ("b200", "tools/libe264bench.so")):
    lib = bench.BenchLib(path)
    for th in (8, 16, 32, 64):
        b = [bufs[i % 16] for i in range(th)]
        lib.run(b, th)
        s, fr, _, d = lib.run(b, th)
        print(name, "threads", th, "fps", round(sum(fr) / s, 1), "per_thread", round(sum(fr) / s / th, 1), flush=True)
PY
fi
cat gpurun_out/host_$TAG.txt; [ -n "$SCALING" ] && cat gpurun_out/scaling_$TAG.txt
