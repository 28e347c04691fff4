#!/bin/bash
# A/B of register budgets (deblock / inter kernels) and hardware queue counts in the 32-stream replay
cd "$(dirname "$0")/.."
for v in "E264B_DBK_MINB=2 E264B_MINB=4" "E264B_DBK_MINB=4 E264B_MINB=4" "E264B_DBK_MINB=4 E264B_MINB=6" "E264B_DBK_MINB=5 E264B_MINB=6" "E264B_DBK_MINB=4 E264B_MINB=8" "E264B_DBK_MINB=5 E264B_MINB=8"; do
  for c in 8 32; do
    env $v CUDA_DEVICE_MAX_CONNECTIONS=$c S=32 STEPS=3 TAG="$v conn$c" python tools/replay_ab.py 2>&1 | tail -1
  done
done
