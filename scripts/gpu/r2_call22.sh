#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2c22_warps_sweep.txt
: > $O
for W in 4 3 2 1; do
  echo "== SIGMA_SCAN_WARPS=$W" >> $O
  SIGMA_SCAN_WARPS=$W timeout 600 python scripts/bench_ss2d_scan.py --images 74 --iters 4 >> $O 2>&1
done
cat $O
