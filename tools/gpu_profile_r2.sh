#!/bin/bash
# Round-2 ncu evidence (one GPU): (1) every launch of ONE bench stream (seed 2000, 60 frames) with duration, DRAM bytes and
# executed instructions -> gpurun_out/ncu_stream_<tag>.csv; (2) --set full of the first pictures (I, P, B, B) ->
# gpurun_out/prof_<tag>.ncu-rep.  Numbers under ncu are cold-cache and serialised: shares and byte counts, not bench values.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-r2}
tools/gen264 -o /tmp/p.264 -W 120 -H 68 -n 60 -s 2000 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 2>/dev/null
timeout -k 5 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/ncu_stream_$TAG.csv tools/b200_decode /tmp/p.264 -q > gpurun_out/ncu_stream_$TAG.log 2>&1
timeout -k 5 400 ncu --set full --clock-control none --import-source on -c 11 -o gpurun_out/prof_$TAG -f tools/b200_decode /tmp/p.264 -q >> gpurun_out/ncu_stream_$TAG.log 2>&1
tail -3 gpurun_out/ncu_stream_$TAG.log; wc -l gpurun_out/ncu_stream_$TAG.csv
