#!/bin/bash
# 2-GPU job: pipelined panel broadcast (GPSS_DIST_PIPE=1) against the single-GPU results, and its timing at n = 50k
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
GPSS_DIST_PIPE=1 GPSS_DIST_PHASES=1 timeout 400 $TR --master-port 29571 scripts/dist_check.py 1100 3000 20000 50000 > gpurun_out/j_dist_pipe.log 2>&1; echo "pipe rc=$?"
grep -E "rep [12]|vs single|rank [0-7] phases|identical|CHECK|sharded" gpurun_out/j_dist_pipe.log | cut -c1-170
