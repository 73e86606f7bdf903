#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block_grads_gpu.py tests/test_modules_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/r2c17_tests.log
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2c17_train_tiny.json 2> gpurun_out/r2c17_train_tiny.err
timeout 600 python scripts/profile_train_step.py > gpurun_out/r2c17_train_profile.txt 2>&1
cat gpurun_out/r2c17_tests.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c17_train_tiny.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('fwd'), d['roofline'].get('bwd'))
PY
head -45 gpurun_out/r2c17_train_profile.txt | cut -c1-160
