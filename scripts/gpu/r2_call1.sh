#!/bin/bash
# round 2, GPU call 1: new parity tests, reference-extension comparison, NPOLY sweep, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2c1_pytest.log
timeout 900 python scripts/bench_vs_ref_ext.py --batch 1 8 --bwd --out gpurun_out/r2c1_ref_ext.json > gpurun_out/r2c1_ref_ext.log 2>&1
for p in 0 1 2 3; do
  SIGMA_SCAN_POLY=$p timeout 300 python scripts/bench_ss2d_scan.py --images 37 --only enc0 enc1 enc2 enc3 > gpurun_out/r2c1_poly$p.log 2>&1
done
SIGMA_SCAN_POLY=2 timeout 300 python -m pytest tests/test_ss2d_scan_gpu.py -q -k fused_scan 2>&1 | tail -15 > gpurun_out/r2c1_poly2_pytest.log
timeout 600 python bench.py > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -5 gpurun_out/r2c1_pytest.log
cat gpurun_out/r2c1_poly*.log
cat gpurun_out/r2c1_bench.json
