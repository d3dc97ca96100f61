#!/bin/bash
# round-2 final profile pass (same layout as r2_call24.sh): default bench line, launch list, ncu --set full of the headline kernels
cd /root/repo
T=r2c37
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log | cut -c1-300
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
NCU="ncu --set full --clock-control none --import-source on -f"
B="--steps 2 --warmup 3 --no-cpu --no-e2e --no-side"
python bench.py $B --block 8192 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches_block8192.csv python bench.py $B --block 8192 > gpurun_out/${T}_list.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel|agc_tmajor|amtail' -s 9 -c 3 -o gpurun_out/prof_${T}_c65536 python bench.py $B > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_c8192 python bench.py $B --channels 8192 > gpurun_out/${T}_ncu2.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel' -s 8 -c 1 -o gpurun_out/prof_${T}_c1 python bench.py $B --channels 1 > gpurun_out/${T}_ncu2c.log 2>&1
ls -la gpurun_out/prof_${T}_*
