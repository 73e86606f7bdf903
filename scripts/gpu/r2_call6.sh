#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SIGMA_PARITY_LOG=$PWD/gpurun_out/r2c6_parity.jsonl
rm -f $SIGMA_PARITY_LOG
timeout 600 python -m pytest tests/test_ss2d_scan_gpu.py -q -k "gemm" 2>&1 | tail -15 > gpurun_out/r2c6_gemm.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_fullsize_golden_gpu.py tests/test_pipeline_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2c6_modules.log
timeout 600 python bench.py --no-by-batch --no-cpu-baseline > gpurun_out/r2c6_bench_x3.json 2> gpurun_out/r2c6_bench_x3.err
timeout 600 python bench.py --no-by-batch --no-cpu-baseline --precision tf32 > gpurun_out/r2c6_bench_tf32.json 2> gpurun_out/r2c6_bench_tf32.err
timeout 600 python scripts/profile_train_step.py > gpurun_out/r2c6_train_profile.txt 2>&1
timeout 900 python bench.py --mode train --steps 5 --warmup 3 --scan-impl ref_ext > gpurun_out/r2c6_train_tiny_refext.json 2> gpurun_out/r2c6_train_tiny_refext.err
timeout 900 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 --scan-impl ref_ext > gpurun_out/r2c6_train_small_bf16_refext.json 2> gpurun_out/r2c6_train_small_bf16_refext.err
timeout 900 python bench.py --mode train --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 > gpurun_out/r2c6_train_small_bf16.json 2> gpurun_out/r2c6_train_small_bf16.err
tail -n 4 gpurun_out/r2c6_gemm.log gpurun_out/r2c6_modules.log; cat $SIGMA_PARITY_LOG
for f in gpurun_out/r2c6_bench_x3.json gpurun_out/r2c6_bench_tf32.json gpurun_out/r2c6_train_*.json; do echo == $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('frac'), d['roofline'].get('fwd'), d['roofline'].get('bwd'), d['config'].get('precision'))
except Exception as e: print('ERR', e)
PY
done
head -60 gpurun_out/r2c6_train_profile.txt | cut -c1-200
tail -n 3 gpurun_out/r2c6_*.err
