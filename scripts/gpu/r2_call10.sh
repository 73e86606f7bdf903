#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_scan_bwd_gpu.py -q -k "widened" 2>&1 | tail -8 > gpurun_out/r2c10_widen.log
timeout 600 python scripts/profile_train_step.py > gpurun_out/r2c10_train_profile.txt 2>&1
timeout 600 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 > gpurun_out/r2c10_train_small_bf16.json 2> gpurun_out/r2c10_train_small_bf16.err
tail -n 4 gpurun_out/r2c10_widen.log
grep -E "sigma::|Self CUDA time|SelectiveScan|FusedSS2D|aten::mm |aten::bmm|aten::copy_|convolution_backward|layer_norm_backward|Optimizer|adamw" gpurun_out/r2c10_train_profile.txt | cut -c1-60,120-200 | head -40
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c10_train_small_bf16.json').read().strip().splitlines()[-1])
print(d['metric'], d['value'], d['ms_per_step'], d['roofline'].get('fwd'), d['roofline'].get('bwd'))
PY
