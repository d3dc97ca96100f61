#!/bin/bash
# round-2 profile pass: default bench line, launch list, and one ncu --set full capture per kernel the bench lines rest on
cd /root/repo
T=r2c24
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.err
NCU="ncu --set full --clock-control none --import-source on -f"
B="--steps 2 --warmup 3 --no-cpu --no-e2e --no-side"
python bench.py $B --block 8192 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches_block8192.csv python bench.py $B --block 8192 > gpurun_out/${T}_list.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel|agc_tmajor|amtail' -s 9 -c 3 -o gpurun_out/prof_${T}_c65536 python bench.py $B > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_c8192 python bench.py $B --channels 8192 > gpurun_out/${T}_ncu2.log 2>&1
LQB_NO_TAILPIPE=1 timeout 600 $NCU -k regex:'agc_tmajor|amtail' -s 6 -c 2 -o gpurun_out/prof_${T}_c8192_tail python bench.py $B --channels 8192 > gpurun_out/${T}_ncu2b.log 2>&1
timeout 600 $NCU -k regex:'lanes_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_c1 python bench.py $B --channels 1 > gpurun_out/${T}_ncu2c.log 2>&1
timeout 600 $NCU -k regex:'fir_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_fir python bench.py --config 2 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu3.log 2>&1
timeout 600 $NCU -k regex:'pipe_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_pipe python bench.py --config 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu4.log 2>&1
timeout 600 $NCU -k regex:'bam_kernel' -s 3 -c 1 -o gpurun_out/prof_${T}_bam python bench.py --next bam --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu5.log 2>&1
timeout 600 $NCU -k regex:'fmst' -s 3 -c 1 -o gpurun_out/prof_${T}_fmst python bench.py --next fmstereo --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu6.log 2>&1
ls -la gpurun_out/prof_${T}_*; tail -2 gpurun_out/${T}_ncu*.log | cut -c1-200
