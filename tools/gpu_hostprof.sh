#!/bin/bash
# Where does a decoder thread's time go on the GPU box?  Parse-only (CPU checker build, reconstruction skipped), the
# reference decoder, and the product library with its host-side time accounting, on one 1080p bench stream.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; TAG=${1:-r2}
tools/gen264 -o /tmp/c2.264 -W 120 -H 68 -n 60 -s 2000 --gop IPB --idr 30 --refs 2 --t8x8 50 --deblock 0 --density 52 --qp 28 --wp 0 2>/dev/null
TIMEFORMAT="%R s wall %U s user %S s sys"
{
echo "== parse only (oracle_decode, E264_NULL_RECON=1)"; for i in 1 2; do time env E264_NULL_RECON=1 E264_HOST_PROFILE=1 oracle/oracle_decode /tmp/c2.264 -q; done
echo "== reference decoder"; for i in 1 2; do time oracle/_ref/ref_decode /tmp/c2.264 -q; done
echo "== product library"; for i in 1 2; do time env E264_HOST_PROFILE=1 tools/b200_decode /tmp/c2.264 -q; done
} > gpurun_out/hostprof_$TAG.txt 2>&1
cat gpurun_out/hostprof_$TAG.txt
