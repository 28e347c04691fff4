#!/bin/bash
# GPU tests + the three bench configs on one box.  Usage: bash tools/gpu_bench_all.sh <tag>
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-r2}
timeout -k 5 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/gputests_$TAG.txt
for c in 1080p 2160p 4320p; do
  timeout -k 5 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_${c}_$TAG.json 2> gpurun_out/bench_${c}_$TAG.err || echo "bench $c failed: $(tail -3 gpurun_out/bench_${c}_$TAG.err)"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${c}_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$c value %.0f fps  e2e %.0f fps (dec_threads %s)  cpu_baseline %s  roofline %s frac %.4f concurrent %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["decoder_n_threads"], d.get("cpu_baseline", {}).get("value"), r["kernel"].split(" ")[0], r["frac"], r["concurrent"]["frac"]))
    print("   per kernel avg us:", {k.split("_")[1]: (round(v["avg_us"]) if v["avg_us"] else None) for k, v in r["per_kernel"].items()}, "launches", d["gpu_launches"])
except Exception as e:
    print("$c: no line", e)
PY
done
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/bench_ref_$TAG.json').read()); print('reference arm', round(d['value']), 'fps', d['cpu_baseline']['cores'], 'cores')"
