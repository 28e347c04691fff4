#!/bin/bash
# A/B of inter-kernel launch shapes (isolated ncu durations + 32-stream replay) and the newest GPU tests
cd "$(dirname "$0")/.."
tools/gen264 -o /tmp/c2.264 -W 120 -H 68 -n 20 -s 2000 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 2>/dev/null
for v in "E264B_MINB=4" "E264B_MINB=6" "E264B_MINB=4 E264B_INTER_WAVES=2" "E264B_MINB=6 E264B_INTER_WAVES=2"; do
  echo "== $v"; env $v ncu --metrics gpu__time_duration.sum --clock-control none -k regex:inter4 -c 12 --csv tools/b200_decode /tmp/c2.264 -q 2>/dev/null | grep inter4 | awk -F, '{gsub(/"/,"",$NF); s+=$NF; n++} END {print n, s/n/1000 " us"}'
done
x=$(oracle/_ref/ref_decode /tmp/c2.264 -q | tail -1); y=$(tools/b200_decode /tmp/c2.264 -q | tail -1); [ "$x" == "$y" ] && echo BITEXACT
S=32 STEPS=3 TAG=static python tools/replay_ab.py 2>&1 | tail -1
E264B_MINB=6 S=32 STEPS=3 TAG=static6 python tools/replay_ab.py 2>&1 | tail -1
python -m pytest tests/test_gen_avc_fixture.py tests/test_kat.py -m gpu -q 2>&1 | tail -3
