#!/bin/bash
# One development step on the GPU box: GPU tests, replay numbers (default, E264B_INTRA_DIV 2 / 4), the three bench configs.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/step_$TAG.txt
run() { echo "== $*" ; env "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
{
timeout -k 5 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run S=32 STEPS=3
run S=32 STEPS=3 E264B_INTRA_DIV=2
run S=32 STEPS=3 E264B_INTRA_DIV=4
run S=32 STEPS=3 E264B_REPLAY_ONLY=1
for c in 1080p 2160p 4320p; do
  timeout -k 5 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_${c}_$TAG.json 2> gpurun_out/bench_${c}_$TAG.err || echo "bench $c failed: $(tail -3 gpurun_out/bench_${c}_$TAG.err)"
  python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_$TAG.json')); print('$c value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value'], 1), d['clocks'], {k.split('_')[1]: round(v['avg_us']) for k, v in d['roofline']['per_kernel'].items() if v['avg_us']})"
done
} 2>&1 | tee $OUT
