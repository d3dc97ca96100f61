#!/bin/bash
TAG=${1:-ab2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
: > gpurun_out/ab_$TAG.jsonl
for args in "--config 3" "--config 4" "--next agc --agc-precision fast" "--next agc" "--config 5" "--next cresamp"; do
  timeout 300 python bench.py $args --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 >> gpurun_out/ab_$TAG.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/ab_$TAG.jsonl"):
    try:
        d = json.loads(l); print("%-10s %10.0f MS/s  %6.3f ms  frac %.3f  %s %s" % (d["metric"].split()[0], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["plan"], d["roofline"].get("segments_ms")))
    except Exception as e:
        print("??", l[:300])
PY
