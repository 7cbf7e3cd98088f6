#!/bin/bash
# 8-GPU job: n = 200 000 objective AND gradient with partitioned storage (BASELINE config 5); then the pipelined panel broadcast at n = 50k
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29561 scripts/part_check.py 200000 > gpurun_out/i_part_n200k_grad.log 2>&1; echo "part rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/i_part_n200k_grad.log | tail -14
for pipe in 0 1; do
  GPSS_DIST_PIPE=$pipe GPSS_DIST_PHASES=1 timeout 200 $TR --master-port 2957$pipe scripts/dist_check.py 50000 > gpurun_out/i_dist_pipe$pipe.log 2>&1; echo "pipe=$pipe rc=$?"
  grep -E "rep [12]|rank [0-7] phases|CHECK" gpurun_out/i_dist_pipe$pipe.log | cut -c1-150
done
