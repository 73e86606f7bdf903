#!/bin/bash
# ncu --set full of the final fused-scan kernels at the benchmarked batch (one launch each) + the fused backward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:ss2d_scan_kernel -s 2 -c 1 -f -o gpurun_out/r02_ncu_ss2d_enc1_b74 python scripts/bench_ss2d_scan.py --images 74 --only enc1 --iters 3 > gpurun_out/r2c18_a.log 2>&1
timeout 600 $NCU -k regex:ss2d_scan_kernel -s 2 -c 1 -f -o gpurun_out/r02_ncu_ss2d_enc0_b74 python scripts/bench_ss2d_scan.py --images 74 --only enc0 --iters 3 > gpurun_out/r2c18_b.log 2>&1
timeout 600 $NCU -k regex:ss2d_scan_kernel -s 2 -c 1 -f -o gpurun_out/r02_ncu_ss2d_dec0_b74 python scripts/bench_ss2d_scan.py --images 74 --only dec0 --iters 3 > gpurun_out/r2c18_c.log 2>&1
timeout 600 $NCU -k regex:ss2d_bwd_kernel -s 1 -c 1 -f -o gpurun_out/r02_ncu_ss2d_bwd_enc1 python scripts/ncu_targets.py fusedbwd > gpurun_out/r2c18_d.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/r2c18_?.log
