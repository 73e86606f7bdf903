#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layernorm_bwd_gpu.py tests/test_block_grads_gpu.py tests/test_modules_gpu.py tests/test_pipeline_gpu.py -q -x 2>&1 | tail -8 > gpurun_out/r2c20_tests.log
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2c20_train_tiny.json 2> gpurun_out/r2c20_train_tiny.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --train-graph > gpurun_out/r2c20_train_tiny_graph.json 2> gpurun_out/r2c20_train_tiny_graph.err
timeout 600 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 > gpurun_out/r2c20_train_small_bf16.json 2> gpurun_out/r2c20_train_small_bf16.err
cat gpurun_out/r2c20_tests.log
for f in gpurun_out/r2c20_train_tiny.json gpurun_out/r2c20_train_tiny_graph.json gpurun_out/r2c20_train_small_bf16.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('fwd'), d['roofline'].get('bwd'))
except Exception as e: print('ERR', e)
PY
done
tail -n 3 gpurun_out/r2c20_train_tiny.err | cut -c1-300
