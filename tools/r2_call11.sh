#!/bin/bash
cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:'lanes_kernel' -s 4 -c 1 -o gpurun_out/prof_r2_lanes_c8192_l2b -f python bench.py --channels 8192 --overlap 0 --steps 2 --warmup 3 --no-cpu --no-e2e --no-side > gpurun_out/r2c11_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'amtail_kernel|agc_tmajor' -s 6 -c 2 -o gpurun_out/prof_r2_tail_c8192 -f python bench.py --channels 8192 --overlap 0 --steps 2 --warmup 3 --no-cpu --no-e2e --no-side > gpurun_out/r2c11_ncu2.log 2>&1
tail -1 gpurun_out/r2c11_ncu2.log | cut -c1-100
