#!/bin/bash
cd /root/repo
T=r2c53
python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
python tools/part_check.py 3000 2>&1 | tail -1
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --channels $C --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/${T}_c${C}_$tag.json 2>&1; }
for C in 1 1024 2048 4096 6144; do run part LQB_X=1; run nopart LQB_NO_PARTITION=1; done
python tools/config1_bench.py --blocks 32 > gpurun_out/${T}_config1.json 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-160
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c53_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), [round(x,3) for x in r.get('segments_ms')], d['gpu']['kernels'][-1][:60])
print(open('gpurun_out/r2c53_config1.json').read()[:330])
PY
