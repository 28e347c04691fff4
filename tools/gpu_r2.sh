#!/bin/bash
# Round-2 GPU session: sanitizer on a small deblocked stream, the validate sweep, then the per-launch device times of one
# 1080p I/P/B clip (ncu launch list, cold-cache and serialised: shares, not absolutes).
# Usage: bash tools/gpu_r2.sh <tag>      -> gpurun_out/{validate,launches,san}_<tag>.*
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-r2}
tools/gen264 -o /tmp/san.264 -W 9 -H 7 -n 6 -s 13 --gop IPB --deblock 0 --t8x8 50 --wp 1 2>/dev/null
timeout -k 5 120 compute-sanitizer --tool memcheck tools/b200_decode /tmp/san.264 -q > gpurun_out/san_$TAG.txt 2>&1; tail -3 gpurun_out/san_$TAG.txt
bash tools/gpu_validate.sh $TAG
tools/gen264 -o /tmp/c2.264 -W 120 -H 68 -n 20 -s 2000 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 2>/dev/null
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv tools/b200_decode /tmp/c2.264 -q > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows: agg[r[4].split("(")[0]].append(float(r[-1]))
for k, v in agg.items(): print("%-40s n=%3d  mean %9.1f us  max %9.1f us" % (k[:40], len(v), sum(v) / len(v) / 1000, max(v) / 1000))
PY
