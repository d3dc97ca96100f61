#!/bin/bash
cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:'amtail8_kernel|agc_tmajor' -s 6 -c 2 -o gpurun_out/prof_r2_tail_c1024 -f python bench.py --channels 1024 --steps 2 --warmup 3 --no-cpu --no-e2e --no-side > gpurun_out/r2c15_ncu.log 2>&1
tail -1 gpurun_out/r2c15_ncu.log | cut -c1-100
