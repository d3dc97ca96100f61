#!/bin/bash
# closing ncu captures on the final tree: headline kernels at the bench launch shape, the front on an eighth of the channels
cd /root/repo
T=r2c49
NCU="ncu --set full --clock-control none --import-source on -f"
B="--steps 2 --warmup 3 --no-cpu --no-e2e --no-side"
timeout 600 $NCU -k regex:'lanes_kernel|agc_tmajor|amtail' -s 9 -c 3 -o gpurun_out/prof_${T}_c65536 python bench.py $B > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_c8192 python bench.py $B --channels 8192 > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/prof_${T}_*
