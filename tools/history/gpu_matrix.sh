#!/bin/bash
# Replay throughput under different concurrency / register settings (each line is one process: the switches are read once).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; OUT=gpurun_out/matrix_${1:-run}.txt
run() { echo "== $*" ; env "$@" timeout -k 5 200 python tools/replay_ab.py 2>&1 | tail -1; }
{
run S=32 STEPS=3
run S=32 STEPS=3
run S=32 STEPS=3 E264B_REPLAY_GRAPH=0 LAUNCH_THREADS=2
run S=32 STEPS=3 E264B_REPLAY_GRAPH=0 LAUNCH_THREADS=1
run S=16 STEPS=3
run S=8 STEPS=3
run S=32 STEPS=3 E264B_MINB=6
run S=32 STEPS=3 E264B_DBK_MINB=4
run S=32 STEPS=3 E264B_MINB=6 E264B_DBK_MINB=4
run S=32 STEPS=3 CUDA_DEVICE_MAX_CONNECTIONS=16
run S=32 STEPS=3 CUDA_DEVICE_MAX_CONNECTIONS=32
run S=32 STEPS=3 CUDA_DEVICE_MAX_CONNECTIONS=4
} 2>&1 | tee $OUT
