#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/r2c4_pytest.log 2>&1; tail -8 gpurun_out/r2c4_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c4_smoke.log 2>&1; tail -3 gpurun_out/r2c4_smoke.log
( time python bench.py ) > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; tail -4 gpurun_out/r2c4_bench.err
for ov in 0 1; do python bench.py --channels 8192 --overlap $ov --no-cpu --no-e2e --no-side > gpurun_out/r2c4_c8192_ov$ov.json 2>&1; done
python - <<'PY'
import json
for f in ['gpurun_out/r2c4_bench.json','gpurun_out/r2c4_c8192_ov0.json','gpurun_out/r2c4_c8192_ov1.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), r.get('frac'), r.get('chain_frac'), r.get('kernel'))
            if d.get('e2e'): print(' e2e', {k:(round(v) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k in ('value','channels_per_gpu','ms_per_step')}, 'i16', round(d['e2e']['int16_iq']['value']), 'pageable', round(d['e2e']['pageable']['value']), d['e2e']['pageable']['h2d_gbs_per_gpu'])
            if d.get('config1'): print(' config1', d['config1'])
            if d.get('side'):
                for k,v in d['side'].items(): print(' ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','frac','plan','kernels','error')})
PY
