#!/bin/bash
# ncu --set full captures of config 4's kernel: the three-warp pipeline (default) and the one-warp seq_kernel (LQB_NO_PIPE=1)
TAG=${1:-c4}
mkdir -p gpurun_out
PROF="python bench.py --config 4 --steps 2 --warmup 3 --no-cpu --no-e2e --block 8192"
timeout 300 $PROF > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'seq_kernel|pipe_kernel' -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_pipe $PROF > gpurun_out/ncu_full_$TAG.log 2>&1
LQB_NO_PIPE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'seq_kernel|pipe_kernel' -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_seq $PROF >> gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
