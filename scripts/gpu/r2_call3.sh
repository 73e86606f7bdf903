#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan_gpu.py tests/test_scan_bwd_gpu.py tests/test_evaluator_gpu.py -q 2>&1 | tail -40 > gpurun_out/r2c3_scan.log
timeout 600 python -m pytest tests/test_scan_grid_gpu.py -q -k "f32" 2>&1 | tail -15 > gpurun_out/r2c3_grid.log
timeout 600 python scripts/bench_op_splits.py > gpurun_out/r2c3_splits_fwd.log 2>&1
timeout 600 python scripts/bench_op_splits.py --bwd --batch 1 8 > gpurun_out/r2c3_splits_bwd.log 2>&1
timeout 900 python scripts/bench_vs_ref_ext.py --batch 8 --dtypes f32 --out gpurun_out/r2c3_ref_ext.json > gpurun_out/r2c3_ref_ext.log 2>&1
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2c3_train_tiny_b2.json 2> gpurun_out/r2c3_train_tiny_b2.err
timeout 900 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 > gpurun_out/r2c3_train_small_bf16.json 2> gpurun_out/r2c3_train_small_bf16.err
tail -n 4 gpurun_out/r2c3_scan.log gpurun_out/r2c3_grid.log
cat gpurun_out/r2c3_splits_fwd.log gpurun_out/r2c3_splits_bwd.log
cat gpurun_out/r2c3_train_tiny_b2.json gpurun_out/r2c3_train_small_bf16.json
tail -n 3 gpurun_out/r2c3_train_tiny_b2.err gpurun_out/r2c3_train_small_bf16.err
