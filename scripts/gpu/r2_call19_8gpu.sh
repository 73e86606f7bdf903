#!/bin/bash
# 8 x B200: the DDP training step (fused core, one gradient bucket) at 1 / 2 / 4 / 8 GPUs — BASELINE config 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --mode train --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2c19_train_tiny_n$n.json 2> gpurun_out/r2c19_train_tiny_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --mode train --gpus $n --steps 5 --warmup 3 > gpurun_out/r2c19_train_tiny_n$n.json 2> gpurun_out/r2c19_train_tiny_n$n.err
  fi
done
for f in gpurun_out/r2c19_*.json; do echo == $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(d['metric'], 'N', d['n_gpus'], d['value'], d['ms_per_step'], d.get('collective'), d['config'].get('peak_mem_gb'))
except Exception as e: print('ERR', e)
PY
done
tail -n 3 gpurun_out/r2c19_*.err | cut -c1-300
