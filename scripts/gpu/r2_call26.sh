#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_ss2d_scan_gpu.py -q -x 2>&1 | tail -1
cd scripts && timeout 900 python bench_ss2d_splits.py > ../gpurun_out/r2c26_split_sweep.txt 2>&1; cd ..
cat gpurun_out/r2c26_split_sweep.txt
