#!/bin/bash
cd /root/repo
LQB_LANES=8 ncu --set full --clock-control none --import-source on -k regex:'lanes_kernel' -s 4 -c 1 -o gpurun_out/prof_r2_lanes_c256_l8b -f python bench.py --channels 256 --overlap 0 --steps 2 --warmup 3 --no-cpu --no-e2e --no-side > gpurun_out/r2c8_ncu1.log 2>&1
tail -1 gpurun_out/r2c8_ncu1.log | cut -c1-100
