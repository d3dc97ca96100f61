#!/bin/bash
cd /root/repo
T=r2c47
run() { tag=$1; shift; env "$@" python bench.py --channels $C --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/${T}_c${C}_$tag.json 2>&1; }
C=65536; run k2 LQB_TAILPIPE_MAX=100000 LQB_TAILPIPE_CHUNKS=2; run k3 LQB_TAILPIPE_MAX=100000 LQB_TAILPIPE_CHUNKS=3; run k4 LQB_TAILPIPE_MAX=100000 LQB_TAILPIPE_CHUNKS=4; run serial LQB_X=1
C=32768; run k2 LQB_TAILPIPE_MAX=100000 LQB_TAILPIPE_CHUNKS=2; run serial LQB_X=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c47_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), [round(x,3) for x in r.get('segments_ms')], d['gpu']['kernels'][2:], d['clocks']['reasons'])
PY
