#!/bin/bash
cd /root/repo
T=r2c43
python -m pytest tests -q -m gpu -x -k "pageable_input" 2>&1 | tail -5
python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for t in 8 4 16; do
LQB_STAGE_THREADS=$t python bench.py --no-cpu --no-side --steps 3 > gpurun_out/${T}_t$t.json 2>&1
done
LQB_NO_STAGING=1 python bench.py --no-cpu --no-side --steps 3 > gpurun_out/${T}_driver.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c43_*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); e=d['e2e']
            print(f, 'c64', round(e['value']), 'i16', round(e['int16_iq']['value']), 'pageable', round(e['pageable']['value']), round(e['pageable']['h2d_gbs_per_gpu'],1), 'GB/s')
PY
