#!/bin/bash
# Luma + chroma deblocking of a band in one 16-warp block (E264B_DBK_MERGED=1) against two 8-warp blocks: GPU parity tests
# under the switch, then replay.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/dbk_merged_$TAG.txt
run() { echo "== $*" ; env "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
{
E264B_DBK_MERGED=1 timeout -k 5 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
run S=32 STEPS=3 E264B_DBK_MERGED=1
run S=32 STEPS=3 E264B_DBK_MERGED=0
run S=32 STEPS=3 E264B_DBK_MERGED=1 E264B_REPLAY_ONLY=4
run S=32 STEPS=3 E264B_DBK_MERGED=1 E264B_REPLAY_ONLY=5
} 2>&1 | tee $OUT
