#!/bin/bash
cd /root/repo
for c in 8192 12288 16384 32768 65536; do
  python bench.py --channels $c --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c22_c${c}.json 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c22_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), d['gpu']['kernels'][1])
PY
