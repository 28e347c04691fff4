#!/bin/bash
# GPU tests, then the replay against the number of streams in flight (and the inter kernel alone), then bench.py.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/inflight_$TAG.txt
run() { echo "== $*" ; env "$@" timeout -k 5 120 python tools/replay_ab.py 2>&1 | grep -E "total" ; }
{
timeout -k 5 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 12 16 20 24 32; do run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=$f; done
run S=32 STEPS=3 E264B_REPLAY_ONLY=1 E264B_REPLAY_INFLIGHT=32
run S=32 STEPS=3 E264B_REPLAY_ONLY=3 E264B_REPLAY_INFLIGHT=32
run S=32 STEPS=3 E264B_REPLAY_ONLY=4 E264B_REPLAY_INFLIGHT=32
run S=32 STEPS=3 E264B_REPLAY_INFLIGHT=32 E264B_DBK_MINB=4
timeout -k 5 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1080p_$TAG.json 2> gpurun_out/bench_1080p_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_1080p_$TAG.json')); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), d['replay'], {k.split('_')[1]: round(v['avg_us']) for k, v in d['roofline']['per_kernel'].items()})"
} 2>&1 | tee $OUT
