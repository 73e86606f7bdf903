#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2c21_enc0_sweep.txt
: > $O
timeout 120 python -m pytest tests/test_layernorm_bwd_gpu.py -q -x 2>&1 | tail -2 >> $O
for W in 3 2 1; do for NST in 0 3 4 8; do
  echo "== WARPS=$W NST=$NST" >> $O
  if [ $NST -eq 0 ]; then SIGMA_SCAN_WARPS=$W SIGMA_SCAN_DEBUG=1 timeout 300 python scripts/bench_ss2d_scan.py --images 74 --only enc0 conmb0 --iters 4 2>&1 | grep -v "^\[ss2d_scan\]" | tail -2 >> $O
  else SIGMA_SCAN_WARPS=$W SIGMA_SCAN_NST=$NST timeout 300 python scripts/bench_ss2d_scan.py --images 74 --only enc0 conmb0 --iters 4 2>&1 | tail -2 >> $O; fi
done; done
cat $O
