#!/bin/bash
# GPU tests (get_frame no longer waits for the device unless the decoder asked for output), e2e with 16 / 20 / 32 application
# threads in both output modes, then bench.py twice.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-run}; OUT=gpurun_out/e2e2_$TAG.txt
{
timeout -k 5 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
cfg = bench.CONFIGS["1080p"]
bufs = bench.generate_streams(cfg, [2000 + i for i in range(32)], 60, "/tmp/e264_bench")
def cpu(): t = os.times(); return t.user + t.system
import subprocess
code = '''
import os, sys, time
sys.path.insert(0, os.getcwd())
import bench
cfg = bench.CONFIGS["1080p"]
bufs = bench.generate_streams(cfg, [2000 + i for i in range(32)], 60, "/tmp/e264_bench")
L = bench.BenchLib(os.path.join(bench.ROOT, sys.argv[1]))
T = int(sys.argv[2])
L.run(bufs, T)
t0 = time.time(); c0 = os.times(); n = 0
for _ in range(3):
    s, fr, _, d = L.run(bufs, T); n += sum(fr)
w = time.time() - t0; c1 = os.times()
print(f"{sys.argv[1]} threads {T} sync_output={os.environ.get('E264_SYNC_OUTPUT','0')}: {n / w:.0f} fps; CPU per frame {1000 * (c1.user + c1.system - c0.user - c0.system) / n:.2f} ms; CPUs busy {(c1.user + c1.system - c0.user - c0.system) / w:.1f}", flush=True)
'''
open("/tmp/e2e_one.py", "w").write(code)
for sync in ("0", "1"):
    for T in (16, 20, 32):
        subprocess.run([sys.executable, "/tmp/e2e_one.py", "tools/libe264bench.so", str(T)], env=dict(os.environ, E264_SYNC_OUTPUT=sync))
for T in (16, 32):
    subprocess.run([sys.executable, "/tmp/e2e_one.py", "oracle/_ref/libe264bench_ref.so", str(T)])
PY
for i in 1 2; do
timeout -k 5 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1080p_${TAG}_$i.json 2> gpurun_out/bench_1080p_${TAG}_$i.err
python -c "
import json; d=json.load(open('gpurun_out/bench_1080p_${TAG}_$i.json')); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), d['replay'], {k.split('_')[1]: round(v['avg_us']) for k, v in d['roofline']['per_kernel'].items()})"
done
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('reference arm', round(d['value']))"
} 2>&1 | tee $OUT
