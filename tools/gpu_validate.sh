#!/bin/bash
# One bounded GPU-box session that answers "is the tree healthy?": guarded decode of a small inter stream (a hang shows in
# 20 s, not in the test suite), the GPU test suite, the 16 bench-stream parity check, one kernel-only replay number.
# Usage (from the repo root on the box):  bash tools/gpu_validate.sh [tag]        -> gpurun_out/validate_<tag>.txt
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/validate_$TAG.txt
{
tools/gen264 -o /tmp/v.264 -W 9 -H 7 -n 8 -s 12 --gop IP --deblock 0 --refs 4 --wp 1 2>/dev/null
a=$(oracle/oracle_decode /tmp/v.264 -q | tail -1); b=$(timeout -k 5 20 tools/b200_decode /tmp/v.264 -q 2>&1 | tail -1)
if [ "$a" != "$b" ]; then echo "GUARD FAILED: port '$a' gpu '$b'"; exit 1; fi
echo "guard ok: $b"
timeout -k 5 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ok=0; for s in $(seq 2000 2015); do
  tools/gen264 -o /tmp/bp.264 -W 120 -H 68 -n 60 -s $s --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 2>/dev/null
  x=$(oracle/_ref/ref_decode /tmp/bp.264 -q | tail -1); y=$(timeout -k 5 60 tools/b200_decode /tmp/bp.264 -q 2>&1 | tail -1)
  [ "$x" == "$y" ] && ok=$((ok+1)) || echo "bench stream $s differs: $x | $y"
done; echo "bench streams bit-exact: $ok of 16"
S=32 STEPS=3 TAG=$TAG timeout -k 5 150 python tools/replay_ab.py 2>&1 | tail -1
} 2>&1 | tee $OUT
