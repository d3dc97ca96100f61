#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu --deselect tests/test_bench_contract.py::test_product_arm_line > gpurun_out/r2c3_pytest.log 2>&1; tail -15 gpurun_out/r2c3_pytest.log
ncu --set full --clock-control none --import-source on -k regex:'lane2_kernel' -s 4 -c 1 -o gpurun_out/prof_r2_lane2a -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --block 8192 > gpurun_out/r2c3_ncu.log 2>&1
tail -3 gpurun_out/r2c3_ncu.log
