#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x > gpurun_out/r2c16_pytest.log 2>&1; tail -4 gpurun_out/r2c16_pytest.log
for c in 2 3; do python bench.py --config $c --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['workload'][:40], round(d['value']), round(d['roofline']['frac'],3), d['config']['plan'])"; done
LQB_FIR_NOUTAP=1 python bench.py --config 2 --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('no-utap', round(d['value']), round(d['roofline']['frac'],3))"
for n in cresamp rfir ssb; do python bench.py --next $n --no-cpu --steps 5 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['workload'][:50], round(d['value']), round(d['roofline']['frac'],3))"; done
