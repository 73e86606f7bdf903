#!/bin/bash
# round 2, GPU call 2: the TMA-staged op-level forward / backward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SIGMA_PARITY_LOG=$PWD/gpurun_out/r2c2_parity.jsonl
rm -f $SIGMA_PARITY_LOG
timeout 600 python -m pytest tests/test_scan_gpu.py -q -x 2>&1 | tail -40 > gpurun_out/r2c2_scan_fwd.log
timeout 600 python -m pytest tests/test_scan_bwd_gpu.py -q 2>&1 | tail -60 > gpurun_out/r2c2_scan_bwd.log
timeout 900 python -m pytest tests/test_ref_ext_gpu.py tests/test_scan_grid_gpu.py -q 2>&1 | tail -60 > gpurun_out/r2c2_ext_grid.log
timeout 900 python scripts/bench_vs_ref_ext.py --batch 1 8 --bwd --out gpurun_out/r2c2_ref_ext.json > gpurun_out/r2c2_ref_ext.log 2>&1
tail -3 gpurun_out/r2c2_scan_fwd.log gpurun_out/r2c2_scan_bwd.log gpurun_out/r2c2_ext_grid.log
tail -2 gpurun_out/r2c2_ref_ext.log | cut -c1-600
