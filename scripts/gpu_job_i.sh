#!/bin/bash
# 8-GPU job: n = 200 000 objective AND gradient with partitioned storage (BASELINE config 5)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29561 scripts/part_check.py 200000 > gpurun_out/i_part_n200k_grad.log 2>&1; echo "part rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/i_part_n200k_grad.log | tail -14
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
