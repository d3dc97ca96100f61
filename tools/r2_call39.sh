#!/bin/bash
cd /root/repo
T=r2c39
python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
run() { tag=$1; shift; env "$@" python bench.py --channels $C --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/${T}_c${C}_$tag.json 2>&1; }
for C in 1 1024 2048 4096 8192 12288; do run stage LQB_X=1; run tile LQB_NO_STAGEBODY=1; done
C=65536; run def LQB_X=1
python tools/config1_bench.py --blocks 32 > gpurun_out/${T}_config1.json 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c39_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), [round(x,3) for x in r.get('segments_ms')], d['gpu']['kernels'][1:3], d['clocks']['reasons'])
print(open('gpurun_out/r2c39_config1.json').read()[:400])
PY
