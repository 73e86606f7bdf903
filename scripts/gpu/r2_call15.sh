#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2c15_npoly_sweep.txt
: > $O
for NP in 0 1 2; do
  echo "== SIGMA_SCAN_NPOLY=$NP" >> $O
  SIGMA_SCAN_NPOLY=$NP timeout 300 python -m pytest tests/test_ss2d_scan_gpu.py -q -x 2>&1 | tail -1 >> $O
  SIGMA_SCAN_NPOLY=$NP timeout 600 python scripts/bench_ss2d_scan.py --images 74 --only enc0 enc1 enc2 enc3 >> $O 2>&1
done
cat $O
