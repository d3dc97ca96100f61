#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x -k "resamp or nco or Resamp or fuzz or golden" > gpurun_out/r2c18_pytest.log 2>&1; tail -2 gpurun_out/r2c18_pytest.log
for sp in 2000 3000; do LQB_PAR_SPAN=$sp python bench.py --config 3 --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('span $sp', round(d['value']), round(d['roofline']['frac'],3), d['config']['plan'])"; done
python bench.py --next cresamp --no-cpu --steps 5 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('cresamp', round(d['value']), round(d['roofline']['frac'],3))"
ncu --set full --clock-control none -k regex:'resamp_par_kernel' -s 3 -c 1 -o gpurun_out/prof_r2_c3 -f python bench.py --config 3 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2c18_ncu.log 2>&1
