#!/bin/bash
# full validation on one B200: GPU test suite, smoke, forward bench (default line), training bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SIGMA_PARITY_LOG=$PWD/gpurun_out/r2c28_parity.jsonl
: > $SIGMA_PARITY_LOG
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2c28_tests.log
unset SIGMA_PARITY_LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c28_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c28_bench_ref.json 2> gpurun_out/r2c28_bench_ref.err
cat gpurun_out/r2c28_tests.log; tail -2 gpurun_out/r2c28_smoke.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c28_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['roofline']['by_dstate'], d['roofline'].get('mufu'), d['clocks'], d.get('gpu_launches'))
print(d.get('other_precision'), d.get('by_batch'))
print(d.get('cpu_baseline'), d.get('gpu_baseline'))
try:
    r=json.loads(open('gpurun_out/r2c28_bench_ref.json').read().strip().splitlines()[-1]); print('REF', r.get('value'), r.get('ms_per_step'), r.get('cpu_baseline'))
except Exception as e: print('REF ERR', e)
PY
tail -n 3 gpurun_out/r2c28_bench.err | cut -c1-300
