#!/bin/bash
# Device timeline of the kernel-only replay (E264B_TRACE): per-kernel spans under load, how many launches of each kind
# overlap, latency of a picture and the gap between a stream's pictures — 32 streams in flight and one stream alone.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}
for S in 32 1; do
  E264B_TRACE=/tmp/trace_$S.csv S=$S STEPS=2 timeout -k 5 200 python tools/replay_ab.py 2>&1 | tail -1
  python tools/trace_report.py /tmp/trace_$S.csv | tee gpurun_out/timeline_S${S}_$TAG.txt
done
