#!/bin/bash
# round 2, call 1: microbenchmarks + baselines of the regimes VERDICT names (few channels, config 1)
set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./tools/ubench_rf > gpurun_out/r2_ubench_rf.log 2>&1
./tools/ubench_ffma2 > gpurun_out/r2_ubench_ffma2.log 2>&1
python tools/config1_bench.py --blocks 64 > gpurun_out/r2_config1_base.json 2> gpurun_out/r2_config1_base.err
for c in 8192 16384 32768; do
  python bench.py --channels $c --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2_base_c$c.json 2> gpurun_out/r2_base_c$c.err
done
python bench.py --channels 8192 --overlap 1 --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2_base_c8192_ov.json 2>&1
tail -n 40 gpurun_out/r2_ubench_rf.log
cat gpurun_out/r2_config1_base.json
