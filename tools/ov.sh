#!/bin/bash
TAG=${1:-ov}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "overlapped" 2>&1 | tail -5 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
: > gpurun_out/ab_$TAG.jsonl
for v in "2 0" "3 1" "2 1"; do
  set -- $v
  LQB_FRONT_RING=$1 timeout 300 python bench.py --overlap $2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 >> gpurun_out/ab_$TAG.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/ab_$TAG.jsonl"):
    try:
        d = json.loads(l); print("%-10s %10.0f MS/s  %6.3f ms  frac %.3f chain %.3f %s %s" % (d["metric"].split()[0], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["chain_frac"], d["config"]["overlap"], d["roofline"].get("segments_ms")))
    except Exception as e:
        print("??", l[:300])
PY
