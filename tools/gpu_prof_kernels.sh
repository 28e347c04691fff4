#!/bin/bash
# ncu --set full with source counters for one launch of each kernel (inter, intra, deblocking, intra rows) of a bench stream,
# then bench.py once.  Usage: bash tools/gpu_prof_kernels.sh <tag>
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}
tools/gen264 -o /tmp/p.264 -W 120 -H 68 -n 8 -s 2000 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 2>/dev/null
timeout -k 5 300 ncu --set full --clock-control none --import-source on --launch-skip 6 --launch-count 4 -o gpurun_out/prof_$TAG -f tools/b200_decode /tmp/p.264 -q > gpurun_out/prof_$TAG.log 2>&1; tail -2 gpurun_out/prof_$TAG.log
timeout -k 5 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1080p_$TAG.json 2> gpurun_out/bench_1080p_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_1080p_$TAG.json')); print('1080p value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'], 1), d['clocks'])"
