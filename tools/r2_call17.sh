#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x -k "resamp or nco or Resamp or fuzz or golden" > gpurun_out/r2c17_pytest.log 2>&1; tail -3 gpurun_out/r2c17_pytest.log
for sp in 1500 2000 3000 4000; do LQB_PAR_SPAN=$sp python bench.py --config 3 --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('span $sp', round(d['value']), round(d['roofline']['frac'],3), d['config']['plan'])"; done
for sp in 2000 3000; do LQB_PAR_SPAN=$sp python bench.py --next cresamp --no-cpu --steps 5 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('cresamp span $sp', round(d['value']), round(d['roofline']['frac'],3))"; done
