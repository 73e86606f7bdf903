#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/tf32_error_budget.py > gpurun_out/r2c5_tf32_budget.log 2>&1
timeout 600 python -m pytest tests/test_block_grads_gpu.py tests/test_pipeline_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2c5_pytest.log
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r2c5_train_tiny.json 2> gpurun_out/r2c5_train_tiny.err
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --train-graph > gpurun_out/r2c5_train_tiny_graph.json 2> gpurun_out/r2c5_train_tiny_graph.err
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --scan-impl ref_ext > gpurun_out/r2c5_train_tiny_refext.json 2> gpurun_out/r2c5_train_tiny_refext.err
timeout 900 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 --train-graph > gpurun_out/r2c5_train_small_bf16_graph.json 2> gpurun_out/r2c5_train_small_bf16_graph.err
timeout 900 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 --scan-impl ref_ext > gpurun_out/r2c5_train_small_bf16_refext.json 2> gpurun_out/r2c5_train_small_bf16_refext.err
timeout 900 python bench.py --model sigma_base --height 720 --width 960 --num-classes 5 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline --no-by-batch > gpurun_out/r2c5_base_720x960_b16.json 2> gpurun_out/r2c5_base.err
cat gpurun_out/r2c5_tf32_budget.log; tail -n 3 gpurun_out/r2c5_pytest.log
for f in gpurun_out/r2c5_train_*.json gpurun_out/r2c5_base_720x960_b16.json; do echo == $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['metric'], d['value'], d['ms_per_step'], d['roofline'].get('fwd'), d['roofline'].get('bwd'), d['roofline'].get('frac'), d['config'].get('cuda_graph'), d['config'].get('peak_mem_gb'))
except Exception as e: print('ERR', e)
PY
done
tail -n 3 gpurun_out/r2c5_*.err
