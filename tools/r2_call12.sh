#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x > gpurun_out/r2c12_pytest.log 2>&1; tail -5 gpurun_out/r2c12_pytest.log
for c in 65536 8192; do
  python bench.py --channels $c --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c12_c$c.json 2> gpurun_out/r2c12_c$c.err
done
LQB_NO_AMTAIL8=1 python bench.py --channels 8192 --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c12_c8192_old.json 2>&1
python tools/config1_bench.py --blocks 32 > gpurun_out/r2c12_config1.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c12_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), d['gpu']['kernels'])
print(open('gpurun_out/r2c12_config1.json').read()[:300])
PY
