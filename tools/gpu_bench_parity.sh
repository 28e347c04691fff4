#!/bin/bash
cd "$(dirname "$0")/.."
for s in $(seq 2000 2015); do
  tools/gen264 -o /tmp/bp.264 -W 120 -H 68 -n 60 -s $s --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 2>/dev/null
  a=$(oracle/_ref/ref_decode /tmp/bp.264 -q | tail -1); b=$(timeout 120 tools/b200_decode /tmp/bp.264 -q 2>&1 | tail -1)
  if [ "$a" == "$b" ]; then echo "OK $s"; else echo "FAIL $s | $a | $b"; oracle/_ref/ref_decode /tmp/bp.264 -c > /tmp/r.txt; tools/b200_decode /tmp/bp.264 -c > /tmp/g.txt 2>&1; diff /tmp/r.txt /tmp/g.txt | head -6; cp /tmp/bp.264 gpurun_out/fail_$s.264; fi
done
