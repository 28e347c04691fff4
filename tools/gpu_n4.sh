cd /root/repo; mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max
timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_N4_r2.json 2> gpurun_out/bench_N4_r2.err; tail -2 gpurun_out/bench_N4_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_N4_r2.json').read().strip().splitlines()[-1]); print('N=4 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'app', d['e2e']['app_threads'], 'dec', d['e2e']['decoder_n_threads'], 'cpus', d['e2e']['usable_cpus'], d['clocks'])"
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/bench_N4_ref_r2.json 2> /dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_N4_ref_r2.json').read().strip().splitlines()[-1]); print('N=4 reference', round(d['value']), d['cpu_baseline']['cores'])"
