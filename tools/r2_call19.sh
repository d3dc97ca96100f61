#!/bin/bash
cd /root/repo
python tools/sanitize_cases.py > gpurun_out/r2_sanitize_plain.log 2>&1; tail -12 gpurun_out/r2_sanitize_plain.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|sanitize cases ok" gpurun_out/r2_sanitize_$tool.log | sort | uniq -c | head -12
done
