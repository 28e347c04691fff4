#!/bin/bash
# What do the small latency-bound kernels cost the inter kernel (whose 16-warp blocks need a whole SM)?  Replay with kernel
# subsets, deblocking blocks limited to one per SM, and block-scheduling priorities per kernel (E264B_PRIO=inter,intra,deblock).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/packing_$TAG.txt
run() { echo "== $*" ; env "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
{
run S=32 STEPS=3
run S=32 STEPS=3 E264B_REPLAY_ONLY=5
run S=32 STEPS=3 E264B_REPLAY_ONLY=3
run S=32 STEPS=3 E264B_DBK_SMEM=100000
run S=32 STEPS=3 E264B_PRIO=0,-1,-1
run S=32 STEPS=3 E264B_PRIO=-1,0,0
run S=32 STEPS=3 E264B_PRIO=0,-2,-1
run S=32 STEPS=3 E264B_PRIO=0,0,-1
} 2>&1 | tee $OUT
