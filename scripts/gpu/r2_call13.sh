#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2c13_ctas_sweep.txt
: > $O
timeout 300 python -m pytest tests/test_ss2d_scan_gpu.py -q -x 2>&1 | tail -2 >> $O
for C in 3 4 5; do
  echo "== SIGMA_SCAN_CTAS=$C" >> $O
  SIGMA_SCAN_CTAS=$C timeout 600 python scripts/bench_ss2d_scan.py --images 74 >> $O 2>&1
done
echo "== SIGMA_SCAN_CTAS=4 parity" >> $O
SIGMA_SCAN_CTAS=4 timeout 300 python -m pytest tests/test_ss2d_scan_gpu.py -q -x 2>&1 | tail -2 >> $O
SIGMA_SCAN_CTAS=5 timeout 300 python -m pytest tests/test_ss2d_scan_gpu.py -q -x 2>&1 | tail -2 >> $O
cat $O
