#!/bin/bash
# quick GPU parity sweep: generator -> reference decoder vs libedge264_b200 (run on a GPU box)
cd "$(dirname "$0")/.."
T=${TMPDIR:-/tmp}/e264par; mkdir -p $T
pass=0; fail=0
run() {  # W H args...
  W=$1; H=$2; shift 2
  timeout 120 tools/gen264 -o $T/x.264 -W $W -H $H "$@" 2>/dev/null || { echo "GENFAIL $W $H $*"; return; }
  a=$(timeout 120 oracle/_ref/ref_decode $T/x.264 -q 2>&1 | tail -1)
  b=$(timeout 120 tools/b200_decode $T/x.264 -q 2>&1 | tail -1)
  if [ "$a" == "$b" ]; then pass=$((pass+1)); echo "OK   ${W}x${H} $*"; else fail=$((fail+1)); echo "FAIL ${W}x${H} $* | ref: $a | gpu: $b"; fi
}
run 4 3 -n 2 -s 1 --gop I --deblock 1 --pcm 0 --t8x8 0
run 9 7 -n 3 -s 7 --gop I --deblock 1 --t8x8 50
run 9 7 -n 3 -s 7 --gop I --deblock 0 --t8x8 50 --pcm 30
run 9 7 -n 3 -s 8 --gop I --deblock 0 --t8x8 50 --scaling 3 --slices 3
run 9 7 -n 3 -s 9 --gop I --deblock 2 --t8x8 50 --slices 4 --cavlc
run 9 7 -n 6 -s 11 --gop IP --deblock 1
run 9 7 -n 6 -s 12 --gop IP --deblock 0 --refs 4 --wp 1
run 9 7 -n 10 -s 13 --gop IPB --deblock 1
run 9 7 -n 10 -s 14 --gop IPB --deblock 0
run 9 7 -n 10 -s 15 --gop IPB --deblock 0 --wp 1
run 9 7 -n 10 -s 16 --gop IPB --deblock 0 --wp 2 --temporal
run 9 7 -n 10 -s 17 --gop IPB --deblock 0 --scaling 3 --refs 4 --mvrange 60 --slices 2
run 120 68 -n 6 -s 2 --gop I --deblock 0
run 120 68 -n 12 -s 3 --gop IPB --deblock 0 --wp 2 --density 25
echo "pass $pass fail $fail"
