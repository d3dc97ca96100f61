#!/bin/bash
# A/B of the single-precision gain loop against the general (double-precision) one: parity first, then throughput.
TAG=${1:-agc}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "agc or config4 or amradio or random_chain or am_" 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
: > gpurun_out/ab_$TAG.jsonl
for env in "" "LQB_AGC_GENERAL=1"; do
  for args in "--config 4" "--next agc" "--config 5"; do
    env $env timeout 300 python bench.py $args --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 >> gpurun_out/ab_$TAG.jsonl
  done
done
python - <<PY
import json
for l in open("gpurun_out/ab_$TAG.jsonl"):
    try:
        d = json.loads(l); print("%-10s %10.0f MS/s  %6.3f ms  frac %.3f  %s %s" % (d["metric"].split()[0], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["plan"], d["roofline"].get("segments_ms")))
    except Exception as e:
        print("??", l[:300])
PY
