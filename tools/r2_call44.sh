#!/bin/bash
# round-2 closing check on one GPU: parity suite, smoke, default bench line
cd /root/repo
T=r2c44
python -m pytest tests -q -m gpu -x > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log | cut -c1-200
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_ref.json 2>&1; tail -1 gpurun_out/${T}_ref.json | cut -c1-400
python - <<'PY'
import json
for l in open('gpurun_out/r2c44_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['segments_ms'], d['roofline']['traffic'], d['clocks'])
        print('e2e', d['e2e']['value'], d['e2e']['int16_iq']['value'], d['e2e']['pageable']['value'])
        print('config1', d['config1']['objects']['msamples_per_s'], d['config1']['chain']['msamples_per_s'])
        for k,v in d['side'].items(): print(k, v['value'], v['frac'])
PY
