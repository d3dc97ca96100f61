#!/bin/bash
cd /root/repo
T=r2c46
python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { tag=$1; shift; env "$@" python bench.py --channels $C --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/${T}_c${C}_$tag.json 2>&1; }
for C in 65536 8192; do run rows LQB_X=1; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c46_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), [round(x,3) for x in r.get('segments_ms')], d['gpu']['kernels'][2:], d['clocks']['reasons'])
PY
