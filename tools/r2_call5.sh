#!/bin/bash
cd /root/repo
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c5_smoke.log 2>&1; tail -2 gpurun_out/r2c5_smoke.log
ncu --set full --clock-control none --import-source on -k regex:'lane2_kernel' -s 4 -c 1 -o gpurun_out/prof_r2_lane2_c8192 -f python bench.py --channels 8192 --overlap 0 --steps 2 --warmup 3 --no-cpu --no-e2e --no-side > gpurun_out/r2c5_ncu.log 2>&1
tail -2 gpurun_out/r2c5_ncu.log | cut -c1-200
