#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SIGMA_PARITY_LOG=$PWD/gpurun_out/r2c4_parity.jsonl
rm -f $SIGMA_PARITY_LOG
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2c4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c4_smoke.log 2>&1
timeout 600 python scripts/bench_op_splits.py > gpurun_out/r2c4_splits_fwd.log 2>&1
timeout 600 python scripts/bench_op_splits.py --bwd --batch 2 8 --splits 0 4 8 16 32 64 > gpurun_out/r2c4_splits_bwd.log 2>&1
timeout 900 python bench.py > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:ss2d_scan_kernel --csv --log-file gpurun_out/r02_scan_traffic.csv python scripts/scan_step_once.py --batch 74 > gpurun_out/r2c4_ncu_traffic.log 2>&1
tail -n 6 gpurun_out/r2c4_pytest.log; cat gpurun_out/r2c4_smoke.log | tail -2
cat gpurun_out/r2c4_splits_fwd.log gpurun_out/r2c4_splits_bwd.log gpurun_out/r2c4_parity.jsonl
cat gpurun_out/r2c4_bench.json; tail -3 gpurun_out/r2c4_bench.err; tail -3 gpurun_out/r2c4_ncu_traffic.log
