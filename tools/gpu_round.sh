#!/bin/bash
# One GPU-box session: parity tests, diagnostics, bench, ncu launch list and one full capture of the
# top kernels.  Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log
timeout 300 python tools/diag_gpu.py > gpurun_out/diag_$TAG.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1
PROF="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --block 8192"
timeout 300 $PROF > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $PROF > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 300 $PROF > gpurun_out/plain2_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'front2_kernel|seq_kernel|amtail_kernel' -s 6 -c 2 -f -o gpurun_out/prof_$TAG $PROF > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 8 gpurun_out/pytest_$TAG.log; cat gpurun_out/diag_$TAG.log; tail -n 2 gpurun_out/bench_$TAG.log; tail -n 3 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out
