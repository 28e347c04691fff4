#!/bin/bash
# Block geometry of the inter kernel (4 / 8 / 16 warps per block): bit-exactness of one bench stream against the reference
# decoder, then the replay numbers (inter alone, whole pipeline) for several numbers of streams in flight.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/variants_$TAG.txt
tools/gen264 -o /tmp/bp.264 -W 120 -H 68 -n 60 -s 2003 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 2>/dev/null
tools/gen264 -o /tmp/wp.264 -W 40 -H 30 -n 12 -s 77 --gop IPB --refs 2 --wp 1 --deblock 0 --density 60 2>/dev/null
{
for v in base w8 w16; do
  if [ $v == base ]; then D=$PWD/edge264_b200; else D=$PWD/edge264_b200/variants/$v; fi
  for f in /tmp/bp.264 /tmp/wp.264; do
    x=$(oracle/_ref/ref_decode $f -q | tail -1); y=$(LD_LIBRARY_PATH=$D timeout -k 5 60 tools/b200_decode $f -q 2>&1 | tail -1)
    [ "$x" == "$y" ] && echo "$v $f bit-exact" || echo "$v $f DIFFERS: $x | $y"
  done
  run() { echo "== $v $*" ; env LD_LIBRARY_PATH=$D E264_LIB_DIR=$D "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
  run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_REPLAY_INFLIGHT=16
  run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=8
  run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=12
  run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=16
done
} 2>&1 | tee $OUT
