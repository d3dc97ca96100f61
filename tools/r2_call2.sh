#!/bin/bash
cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/r2c2_pytest.log 2>&1; tail -5 gpurun_out/r2c2_pytest.log
for c in 65536 8192; do
  python bench.py --channels $c --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2c2_lanes_c$c.json 2> gpurun_out/r2c2_lanes_c$c.err
  LQB_NO_LANES=1 python bench.py --channels $c --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2c2_old_c$c.json 2>&1
done
python tools/config1_bench.py --blocks 64 > gpurun_out/r2c2_config1.json 2> gpurun_out/r2c2_config1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c2_*_c*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), r.get('frac'))
PY
cat gpurun_out/r2c2_config1.json; tail -3 gpurun_out/r2c2_lanes_c65536.err
