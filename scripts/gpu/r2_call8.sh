#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export SIGMA_PARITY_LOG=$PWD/gpurun_out/r2c8_parity.jsonl
rm -f $SIGMA_PARITY_LOG
timeout 600 python -m pytest tests/test_ss2d_scan_gpu.py -q -k "conv3x3 or gemm" 2>&1 | tail -12 > gpurun_out/r2c8_gemm_conv.log
SIGMA_X3_KEEP_HI=0 timeout 600 python -m pytest tests/test_ss2d_scan_gpu.py -q -k "conv3x3 or gemm" 2>&1 | tail -12 > gpurun_out/r2c8_gemm_conv_nohi.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_fullsize_golden_gpu.py tests/test_evaluator_gpu.py -q 2>&1 | tail -12 > gpurun_out/r2c8_modules.log
timeout 600 python bench.py > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
SIGMA_X3_KEEP_HI=0 timeout 600 python bench.py --no-by-batch --no-cpu-baseline > gpurun_out/r2c8_bench_nohi.json 2> gpurun_out/r2c8_bench_nohi.err
timeout 1200 python scripts/bench_vs_ref_ext.py --batch 1 8 --bwd --out gpurun_out/r2c8_ref_ext.json > gpurun_out/r2c8_ref_ext.log 2>&1
for w in opfwd opfwd_n4 opbwd gemm conv; do
  case $w in
    opfwd|opfwd_n4) k="regex:scan_op_tma_kernel";;
    opbwd) k="regex:scan_op_bwd_tma_kernel";;
    gemm|conv) k="regex:gemm_tf32_kernel";;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k $k -c 4 -f -o gpurun_out/r02_ncu_$w python scripts/ncu_targets.py $w > gpurun_out/r2c8_ncu_$w.log 2>&1
done
tail -n 4 gpurun_out/r2c8_gemm_conv.log gpurun_out/r2c8_gemm_conv_nohi.log gpurun_out/r2c8_modules.log
cat $SIGMA_PARITY_LOG | grep -E "tiny|b2"
for f in gpurun_out/r2c8_bench.json gpurun_out/r2c8_bench_nohi.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'].get('frac'), d.get('other_precision'), d.get('by_batch'))
except Exception as e: print('ERR', e)
PY
done
tail -n 2 gpurun_out/r2c8_ncu_*.log gpurun_out/r2c8_bench.err | cut -c1-200
ls -la gpurun_out/*.ncu-rep
