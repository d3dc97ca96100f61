#!/bin/bash
# Device-resident throughput of the SURVEY 8(f) rows (one JSON line each) -> gpurun_out/next_<tag>.jsonl
TAG=${1:-r1}
mkdir -p gpurun_out
: > gpurun_out/next_$TAG.jsonl
for row in bam ssb fmstereo rrrf cresamp rfir; do
  timeout 300 python bench.py --next $row --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 >> gpurun_out/next_$TAG.jsonl
done
for row in iir nco "agc" "agc --agc-precision fast" fm deemph resamp; do
  timeout 300 python bench.py --next $row --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 >> gpurun_out/next_$TAG.jsonl
done
for c in 2 3 4; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 >> gpurun_out/next_$TAG.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/next_$TAG.jsonl"):
    try:
        d = json.loads(l); print("%-10s %10.0f MS/s  %6.3f ms  frac %.3f  %s" % (d["metric"].split()[0], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["plan"]))
    except Exception as e:
        print("??", l[:200])
PY
