#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ss2d_scan_gpu.py tests/test_fullsize_golden_gpu.py tests/test_modules_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r2c12_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-by-batch > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
cat gpurun_out/r2c12_tests.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c12_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['by_dstate'], d['clocks'])
PY
tail -n 3 gpurun_out/r2c12_bench.err | cut -c1-300
