#!/bin/bash
# What the driver does at the end of a round, on the committed tree: GPU tests, smoke(), bench.py with its defaults, the reference arm.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}
{
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 600 python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default_$TAG.json')); print('bench.py defaults: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu_baseline', round(d['cpu_baseline']['value']), 'steps', d['steps'], 'launches', d['gpu_launches'], d['clocks'])"
timeout -k 5 600 python bench.py --impl reference > gpurun_out/bench_default_ref_$TAG.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_default_ref_$TAG.json')); print('reference arm defaults:', round(d['value']))"
} 2>&1 | tee gpurun_out/final_check_$TAG.txt
