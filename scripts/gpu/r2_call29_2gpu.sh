#!/bin/bash
# 2 x B200: the driver's multi-GPU launch of both bench arms (forward replicas; reference arm on rank 0 only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2c29_bench_n2.json 2> gpurun_out/r2c29_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29778 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2c29_ref_n2.json 2> gpurun_out/r2c29_ref_n2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2c29_bench_n2.json','gpurun_out/r2c29_ref_n2.json'):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d.get('impl'), d['n_gpus'], d['value'], d['ms_per_step'], d.get('e2e'), (d.get('roofline') or {}).get('frac'), d.get('scaling'))
    except Exception as e: print(f,'ERR',e)
PY
tail -3 gpurun_out/r2c29_bench_n2.err | cut -c1-300
