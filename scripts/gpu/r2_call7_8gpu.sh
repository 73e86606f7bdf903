#!/bin/bash
# 8 x B200: DDP training step at 1 / 2 / 4 / 8 GPUs (BASELINE config 4: Sigma-tiny, 2 images per GPU, NCCL gradient all-reduce),
# Sigma-base 720x960 forward on 8 replicas (config 5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2c7_smi.txt
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --mode train --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2c7_train_tiny_n$n.json 2> gpurun_out/r2c7_train_tiny_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --mode train --gpus $n --steps 5 --warmup 3 > gpurun_out/r2c7_train_tiny_n$n.json 2> gpurun_out/r2c7_train_tiny_n$n.err
  fi
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --model sigma_base --height 720 --width 960 --num-classes 5 --batch 32 --steps 5 --warmup 3 --no-cpu-baseline --no-by-batch > gpurun_out/r2c7_base_720x960_n8.json 2> gpurun_out/r2c7_base_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --mode train --gpus 8 --model sigma_small --num-classes 40 --amp bf16 --steps 5 --warmup 3 > gpurun_out/r2c7_train_small_bf16_n8.json 2> gpurun_out/r2c7_train_small_bf16_n8.err
for f in gpurun_out/r2c7_*.json; do echo == $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(d['metric'], 'N', d['n_gpus'], d['value'], d['ms_per_step'], d.get('collective'), d['config'].get('peak_mem_gb'))
except Exception as e: print('ERR', e)
PY
done
tail -n 4 gpurun_out/r2c7_*.err | cut -c1-300
