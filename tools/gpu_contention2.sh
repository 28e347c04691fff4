#!/bin/bash
# Follow-up of gpu_contention.sh: the inter kernel alone against the number of streams in flight and the register budget,
# then the whole pipeline; e2e against application threads; GPU tests (new deblocking hand-over).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/contention2_$TAG.txt
run() { echo "== $*" ; env "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
{
timeout -k 5 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 1 2 4 8 16 32; do run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_REPLAY_INFLIGHT=$f; done
run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_MINB=6
run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_MINB=8
run S=32 STEPS=3 E264B_REPLAY_ONLY=5
run S=32 STEPS=3 E264B_REPLAY_ONLY=5 E264B_REPLAY_INFLIGHT=32
run S=32 STEPS=3
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=8
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=32
run S=32 STEPS=3 E264B_REPLAY_ONLY=4
echo "== e2e"; timeout -k 5 300 python tools/e2e_ab.py 2>&1 | grep -E "fps"
bash tools/gpu_hostprof.sh $TAG > /dev/null 2>&1; grep -E "parse_slice|wall" gpurun_out/hostprof_$TAG.txt
} 2>&1 | tee $OUT
