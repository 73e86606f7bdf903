#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ss2d_scan_gpu.py tests/test_modules_gpu.py tests/test_fullsize_golden_gpu.py -q -x 2>&1 | tail -2
SIGMA_GEMM_BN_RULE=old SIGMA_SCAN_SPLIT_RULE=old timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c27_bench_oldrules.json 2> gpurun_out/r2c27_old.err
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c27_bench_new.json 2> gpurun_out/r2c27_new.err
for f in gpurun_out/r2c27_bench_oldrules.json gpurun_out/r2c27_bench_new.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['by_batch'].items()})
PY
done
tail -2 gpurun_out/r2c27_new.err | cut -c1-200
