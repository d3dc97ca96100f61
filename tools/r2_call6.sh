#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu -x --deselect tests/test_parity_gpu.py::test_fused_chain_matches_stagewise > gpurun_out/r2c6_pytest.log 2>&1; tail -6 gpurun_out/r2c6_pytest.log
for c in 65536 8192; do
  python bench.py --channels $c --overlap 0 --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c6_c$c.json 2> gpurun_out/r2c6_c$c.err
done
for l in 2 4 8; do LQB_LANES=$l python tools/config1_bench.py --blocks 32 > gpurun_out/r2c6_config1_l$l.json 2> gpurun_out/r2c6_config1_l$l.err; done
for c in 256 1024 2048 4096; do for l in 2 4 8; do
  LQB_LANES=$l python bench.py --channels $c --overlap 0 --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c6_c${c}_l$l.json 2>&1
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c6_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), r.get('kernel'))
for f in sorted(glob.glob('gpurun_out/r2c6_config1_l*.json')):
    print(f, open(f).read()[:400])
PY
