cd /root/repo; mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_N2_r2.json 2> gpurun_out/bench_N2_r2.err; tail -3 gpurun_out/bench_N2_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_N2_r2.json').read().strip().splitlines()[-1]); print('N=2 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['app_threads'], d['e2e']['decoder_n_threads'], d['e2e']['usable_cpus'], d['clocks'])"
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_N2_ref_r2.json 2> gpurun_out/bench_N2_ref_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_N2_ref_r2.json').read().strip().splitlines()[-1]); print('N=2 reference', round(d['value']), d['cpu_baseline']['cores'])"
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 4320p --steps 3 --warmup 3 > gpurun_out/bench_N2_4320p_r2.json 2> gpurun_out/bench_N2_4320p_r2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_N2_4320p_r2.json').read().strip().splitlines()[-1]); print('N=2 4320p value', round(d['value']), 'e2e', round(d['e2e']['value'], 1))"
