#!/bin/bash
cd /root/repo
T=r2c52
python tools/part_check.py 300 2>&1 | tail -3
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --channels $C --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/${T}_c${C}_$tag.json 2>&1; }
C=8192
run part4 LQB_TIMEPIPE_PARTITION=1 LQB_TIMEPIPE_MAX=10000 LQB_TIMEPIPE_SLICES=4
run part6 LQB_TIMEPIPE_PARTITION=1 LQB_TIMEPIPE_MAX=10000 LQB_TIMEPIPE_SLICES=6
run part8 LQB_TIMEPIPE_PARTITION=1 LQB_TIMEPIPE_MAX=10000 LQB_TIMEPIPE_SLICES=8
C=4096
run part6 LQB_TIMEPIPE_PARTITION=1 LQB_TIMEPIPE_MAX=10000 LQB_TIMEPIPE_SLICES=6
run def LQB_X=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c52_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), [round(x,3) for x in r.get('segments_ms')], d['gpu']['kernels'][-1])
PY
