#!/bin/bash
# compute-sanitizer over small streams that reach every kernel (I picture: rows kernel; P/B with intra macroblocks: inter,
# intra, deblocking; 8x8 transform, weighted prediction): memcheck, racecheck (shared-memory hazards), synccheck.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/sanitize_$TAG.txt
tools/gen264 -o /tmp/san1.264 -W 9 -H 7 -n 6 -s 13 --gop IPB --deblock 0 --t8x8 50 --wp 1 --intra-pct 15 2>/dev/null
tools/gen264 -o /tmp/san2.264 -W 20 -H 18 -n 4 -s 5 --gop IPB --deblock 0 --refs 2 --density 60 2>/dev/null
{
for f in /tmp/san1.264 /tmp/san2.264; do
  want=$(oracle/_ref/ref_decode $f -q | tail -1)
  for tool in memcheck racecheck synccheck; do
    timeout -k 5 280 compute-sanitizer --tool $tool tools/b200_decode $f -q > /tmp/san.log 2>&1
    got=$(grep "^frames" /tmp/san.log | tail -1); sum=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" /tmp/san.log | tail -1)
    echo "$f $tool: $sum | output $([ "$got" == "$want" ] && echo bit-exact || echo "DIFFERS ($got vs $want)")"
    grep -E "Error|hazard|Race" /tmp/san.log | head -5
  done
done
} 2>&1 | tee $OUT
