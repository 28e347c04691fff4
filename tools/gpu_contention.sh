#!/bin/bash
# Why do 16 pictures in flight stretch every kernel 4x?  Replay throughput with subsets of the kernels, different numbers of
# streams in flight, and the deblocking blocks limited to one per SM; block placement per SM (E264B_DIAG).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; OUT=gpurun_out/contention_${1:-run}.txt
run() { echo "== $*" ; env "$@" E264B_DIAG=1 timeout -k 5 200 python tools/replay_ab.py 2>&1 | grep -E "diag|total" ; }
{
run S=32 STEPS=3
run S=32 STEPS=3 E264B_REPLAY_ONLY=4
run S=32 STEPS=3 E264B_REPLAY_ONLY=4 E264B_REPLAY_INFLIGHT=4
run S=32 STEPS=3 E264B_REPLAY_ONLY=4 E264B_REPLAY_INFLIGHT=1
run S=32 STEPS=3 E264B_REPLAY_ONLY=4 E264B_REPLAY_INFLIGHT=32
run S=32 STEPS=3 E264B_REPLAY_ONLY=3
run S=32 STEPS=3 E264B_REPLAY_ONLY=1
run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_REPLAY_INFLIGHT=1
run S=32 STEPS=3 E264B_REPLAY_ONLY=2
run S=32 STEPS=3 E264B_DBK_SMEM=100000
run S=32 STEPS=3 E264B_DBK_SMEM=100000 E264B_REPLAY_ONLY=4
run S=32 STEPS=3 E264B_DBK_SMEM=100000 E264B_REPLAY_INFLIGHT=32
} 2>&1 | tee $OUT
