#!/bin/bash
cd /root/repo
T=r2c25
python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
for c in 65536 8192 1024 1; do
  python bench.py --channels $c --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/${T}_c${c}.json 2>&1
done
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'lanes_kernel' -s 4 -c 1 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-side 2>/dev/null | grep lanes_kernel | awk -F'","' '{print $(NF-2), $NF}'
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c25_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), round(r.get('frac'),3), d['gpu']['kernels'][1:])
PY
