#!/bin/bash
# Third contention session: inter kernel with one copy of the window loads, intra kernel with overlapped flag waits,
# blocks per intra launch (E264B_INTRA_DIV), streams in flight.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/contention3_$TAG.txt
run() { echo "== $*" ; env "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
{
timeout -k 5 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run S=32 STEPS=3
run S=32 STEPS=3 E264B_REPLAY_ONLY=1
run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_REPLAY_INFLIGHT=1
run S=32 STEPS=3 E264B_REPLAY_ONLY=3
run S=32 STEPS=3 E264B_INTRA_DIV=2
run S=32 STEPS=3 E264B_INTRA_DIV=4
run S=32 STEPS=3 E264B_INTRA_DIV=8
run S=32 STEPS=3 E264B_INTRA_DIV=4 E264B_REPLAY_ONLY=3
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=8
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=12
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=12 E264B_INTRA_DIV=4
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=24 E264B_INTRA_DIV=4
} 2>&1 | tee $OUT
