#!/bin/bash
# Block geometry of the P/B intra kernel (INTRA_WARPS 16 default / variants 4, 8) x list entries per warp (E264B_INTRA_DIV).
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/intra_geometry_$TAG.txt
tools/gen264 -o /tmp/bp.264 -W 120 -H 68 -n 30 -s 2003 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 --intra-pct 10 2>/dev/null
{
timeout -k 5 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in base iw4 iw8; do
  if [ $v == base ]; then D=$PWD/edge264_b200; else D=$PWD/edge264_b200/variants/$v; fi
  x=$(oracle/_ref/ref_decode /tmp/bp.264 -q | tail -1); y=$(LD_LIBRARY_PATH=$D timeout -k 5 60 tools/b200_decode /tmp/bp.264 -q 2>&1 | tail -1)
  [ "$x" == "$y" ] && echo "$v bit-exact" || echo "$v DIFFERS: $x | $y"
  run() { echo "== $v $*" ; env LD_LIBRARY_PATH=$D E264_LIB_DIR=$D "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
  run S=32 STEPS=3 E264B_INTRA_DIV=1
  run S=32 STEPS=3 E264B_INTRA_DIV=2
  run S=32 STEPS=3 E264B_INTRA_DIV=4
done
} 2>&1 | tee $OUT
