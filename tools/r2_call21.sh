#!/bin/bash
cd /root/repo
for c in 16384 32768; do for l in 2 4; do
  LQB_LANES=$l python bench.py --channels $c --no-cpu --no-e2e --no-side --steps 10 > gpurun_out/r2c21_c${c}_l$l.json 2>&1
done; done
LQB_LANES=4 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__warps_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio --clock-control none -k regex:'lanes_kernel' -s 4 -c 1 --csv python bench.py --channels 8192 --steps 2 --warmup 3 --no-cpu --no-e2e --no-side 2>/dev/null | grep lanes_kernel | awk -F'","' '{print $(NF-2), $NF}' 
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c21_c*.json')):
    for l in open(f):
        if l.startswith('{"metric'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f, round(d['value']), 'MS/s', round(d['ms_per_step'],3), r.get('segments_ms'), d['gpu']['kernels'][1])
PY
