#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x > gpurun_out/r2c23_pytest.log 2>&1; tail -4 gpurun_out/r2c23_pytest.log
for c in 1 1024 8192 16384; do
  python bench.py --channels $c --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c23_c${c}.json 2>&1
  LQB_NO_TAILPIPE=1 python bench.py --channels $c --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c23_c${c}_nopipe.json 2>&1
done
python tools/config1_bench.py --blocks 32 > gpurun_out/r2c23_config1.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c23_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), d['gpu']['kernels'][2:])
print(open('gpurun_out/r2c23_config1.json').read()[:300])
PY
