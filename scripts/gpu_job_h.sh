#!/bin/bash
# 2-GPU job: partitioned-storage gradient against the single-GPU gradient
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 scripts/part_check.py 700 3000 20000 > gpurun_out/h_part_grad.log 2>&1; echo "part rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/h_part_grad.log | tail -40
