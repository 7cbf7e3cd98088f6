#!/bin/bash
# ROUND 2, GPU call (2 GPUs): predictive variance on partitioned handles (variance_partitioned) against one GPU, with the rest of part_check.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 scripts/part_check.py 3000 20000 > $O/r2p_part_check.log 2>&1; echo "part_check rc=$?"; grep -v '^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$\|   g = \|ref g' $O/r2p_part_check.log | tail -30
