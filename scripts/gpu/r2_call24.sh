#!/bin/bash
# one eager step of the benchmark model (B = 74, default precision) under ncu: launch list (time share per kernel) + DRAM bytes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r02_step_b74_x3.csv python scripts/scan_step_once.py --batch 74 > gpurun_out/r2c24.log 2>&1
tail -2 gpurun_out/r2c24.log
python scripts/summarize_launches.py gpurun_out/r02_step_b74_x3.csv | head -30
python scripts/ncu_scan_traffic.py gpurun_out/r02_step_b74_x3.csv --batch 74 --out gpurun_out/r02_scan_traffic_final.json
