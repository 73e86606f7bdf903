#!/bin/bash
# B = 1 latency: per-kernel durations of one eager forward (ncu, time only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_step_b1_x3.csv python scripts/scan_step_once.py --batch 1 > gpurun_out/r2c25.log 2>&1
tail -1 gpurun_out/r2c25.log
python scripts/summarize_launches.py gpurun_out/r02_step_b1_x3.csv | head -32
