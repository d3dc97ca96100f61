#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x > gpurun_out/r2c10_pytest.log 2>&1; tail -3 gpurun_out/r2c10_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
for c in 65536 8192; do
  python bench.py --channels $c --overlap 0 --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c10_c$c.json 2> gpurun_out/r2c10_c$c.err
done
python bench.py --channels 8192 --overlap 1 --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c10_c8192_ov.json 2>&1
LQB_LANES=4 python bench.py --channels 8192 --overlap 0 --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c10_c8192_l4.json 2>&1
python tools/config1_bench.py --blocks 32 > gpurun_out/r2c10_config1.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c10_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), r.get('kernel'))
print(open('gpurun_out/r2c10_config1.json').read()[:300])
PY
