#!/bin/bash
# Round-2 evidence session (one GPU): GPU tests + the three bench configs + the reference arm, the host-side time
# accounting, then the ncu launch list / DRAM bytes of one bench stream and a --set full capture of its first pictures.
# Usage: bash tools/gpu_r2_evidence.sh <tag>      -> gpurun_out/*_<tag>.*
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-r2}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/host_$TAG.txt; nproc >> gpurun_out/host_$TAG.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/host_$TAG.txt
bash tools/gpu_bench_all.sh $TAG
bash tools/gpu_hostprof.sh $TAG > /dev/null 2>&1; tail -20 gpurun_out/hostprof_$TAG.txt
bash tools/gpu_profile_r2.sh $TAG
