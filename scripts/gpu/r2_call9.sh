#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_bwd_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r2c9_fused_bwd.log
timeout 300 python -m pytest tests/test_block_grads_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2c9_block_grads.log
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2c9_train_tiny.json 2> gpurun_out/r2c9_train_tiny.err
timeout 600 python bench.py --mode train --steps 5 --warmup 3 --train-graph > gpurun_out/r2c9_train_tiny_graph.json 2> gpurun_out/r2c9_train_tiny_graph.err
timeout 300 python -m pytest tests/test_scan_gpu.py tests/test_scan_bwd_gpu.py -q 2>&1 | tail -8 > gpurun_out/r2c9_scan.log
timeout 600 python bench.py --no-by-batch --no-cpu-baseline > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
tail -n 25 gpurun_out/r2c9_fused_bwd.log; tail -n 6 gpurun_out/r2c9_block_grads.log gpurun_out/r2c9_scan.log
for f in gpurun_out/r2c9_train_tiny.json gpurun_out/r2c9_train_tiny_graph.json gpurun_out/r2c9_bench.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('frac'), d['roofline'].get('fwd'), d['roofline'].get('bwd'), d['config'].get('cuda_graph'))
except Exception as e: print('ERR', e)
PY
done
tail -n 5 gpurun_out/r2c9_train_tiny.err gpurun_out/r2c9_bench.err | cut -c1-300
